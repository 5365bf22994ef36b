"""GPU parity tests (run on a B200 via `pytest -m gpu`): the CUDA path through the C ABI vs the CPU oracle."""
import itertools

import numpy as np
import pytest

from npairloss_b200 import capi, synth

pytestmark = pytest.mark.gpu

PRECS = [capi.PREC_FP32_FP16X2, capi.PREC_FP32_BF16X3, capi.PREC_BF16]
PREC_NAME = {0: "bf16x3", 1: "bf16", 2: "fp16x2"}


@pytest.fixture(scope="module")
def cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    assert torch.cuda.get_device_capability(0)[0] == 10
    return torch


@pytest.mark.parametrize("backend", [capi.GEMM_SIMT_CHECK, capi.GEMM_TCGEN05])
@pytest.mark.parametrize("prec", PRECS)
def test_split_gemm_vs_fp64(cuda, prec, backend):
    torch = cuda
    tol = {capi.PREC_FP32_BF16X3: 3e-6, capi.PREC_FP32_FP16X2: 3e-6, capi.PREC_BF16: 3e-2}[prec]
    g = torch.Generator(device="cpu").manual_seed(5)
    for (M, Nn, K) in [(128, 256, 64), (128, 256, 128), (256, 512, 512), (120, 120, 100), (1, 7, 3), (300, 1000, 1024), (129, 257, 65)]:
        A = torch.randn(M, K, generator=g)
        B = torch.randn(Nn, K, generator=g)
        A = A / A.norm(dim=1, keepdim=True)
        B = B / B.norm(dim=1, keepdim=True)
        C = capi.debug_gemm(prec, backend, A.cuda(), B.cuda()).cpu()
        ref = (A.double() @ B.double().T)
        err = (C.double() - ref).abs().max().item()
        assert err <= tol, f"prec={PREC_NAME[prec]} backend={backend} shape={(M, Nn, K)} err={err:.3e}"


def test_split_gemm_large_dynamic_range(cuda):
    """bf16x3 keeps fp32 accuracy for un-normalised operands; fp16x2 relies on the power-of-two pre-scale."""
    torch = cuda
    g = torch.Generator(device="cpu").manual_seed(9)
    A = torch.randn(256, 256, generator=g) * 37.0
    B = torch.randn(512, 256, generator=g) * 37.0
    ref = A.double() @ B.double().T
    scale = ref.abs().max().item()
    for prec in (capi.PREC_FP32_BF16X3, capi.PREC_FP32_FP16X2):
        C = capi.debug_gemm(prec, capi.GEMM_TCGEN05, A.cuda(), B.cuda()).cpu().double()
        assert (C - ref).abs().max().item() <= 4e-6 * scale


@pytest.mark.parametrize("backend", [capi.GEMM_SIMT_CHECK, capi.GEMM_TCGEN05])
def test_kat_survey_9_3(cuda, oracle, backend):
    from gpu_harness import gpu_step_world
    x = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=np.float32)
    lab = np.array([0, 0, 1, 1], dtype=np.float32)
    g = gpu_step_world(x, lab, 4, 1, synth.DEFAULT_MINING, capi.PREC_FP32_BF16X3, backend)
    assert g["tops"][0, 0] == pytest.approx(0.5514447139, rel=1e-6)
    assert list(g["tops"][0, 1:4]) == [1.0, 1.0, 1.0]
    assert g["tops"][0, 4] == pytest.approx(1.0, rel=1e-6)
    gg = 0.1059707788
    np.testing.assert_allclose(g["dx"], np.array([[-gg, gg], [-gg, gg], [gg, -gg], [gg, -gg]], dtype=np.float32), rtol=1e-5)


@pytest.mark.parametrize("backend", [capi.GEMM_SIMT_CHECK, capi.GEMM_TCGEN05])
@pytest.mark.parametrize("world,bwd_exchange", [(1, 0), (2, 0), (2, 1), (3, 0)])
def test_all_mining_modes_small(cuda, oracle, world, bwd_exchange, backend):
    """Every (region, method) combination, both GEMM engines, emulated ranks, both backward exchange forms
    (bwd_exchange 0 = row-scalar exchange over the bitwise-symmetric GEMM, 1 = the reference's reduce-scatter form)."""
    from gpu_harness import check_parity
    Q, D = 48, 40
    x, lab = synth.make_inputs(Q * world, D, seed=100 + world, imgs_per_class=3, noise=0.7)
    for apR, apM, anR, anM in itertools.product([0, 1], range(5), [0, 1], range(5)):
        mining = dict(margin_ident=0.02, margin_diff=-0.03, identsn=-0.4, diffsn=-0.3,
                      ap_region=apR, ap_method=apM, an_region=anR, an_method=anM)
        check_parity(oracle, x, lab, Q, world, mining, capi.PREC_FP32_FP16X2, backend, loss_weight=0.7,
                     tag=f"w{world} x{bwd_exchange} b{backend} {apR}{apM}{anR}{anM}", bwd_exchange=bwd_exchange)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("bwd_exchange", [0, 1])
def test_multirank_precisions_and_exchange_forms(cuda, oracle, prec, bwd_exchange):
    """world = 4 at a size with several tiles per rank; bf16x3 always takes the reduce-scatter form."""
    from gpu_harness import check_parity
    Q, world, D = 160, 4, 200
    x, lab = synth.make_inputs(Q * world, D, seed=77, imgs_per_class=2, noise=2.5)
    r = check_parity(oracle, x, lab, Q, world, synth.USAGE_MINING, prec, capi.GEMM_TCGEN05, tag=f"w4 {PREC_NAME[prec]} x{bwd_exchange}",
                     bwd_exchange=bwd_exchange)
    print(PREC_NAME[prec], bwd_exchange, r)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name", ["C1", "C2"])
def test_baseline_configs_small(cuda, oracle, name, prec):
    from gpu_harness import check_parity
    c = synth.CONFIGS[name]
    x, lab = synth.config_inputs(name)
    r = check_parity(oracle, x, lab, c["B"], 1, c["mining"], prec, capi.GEMM_TCGEN05, tag=f"{name} {PREC_NAME[prec]}")
    print(name, PREC_NAME[prec], r)


@pytest.mark.parametrize("mining_name", ["usage", "default", "local_rel"])
@pytest.mark.parametrize("shape", [(120, 1, 1024), (30, 2, 64), (256, 4, 128), (1000, 1, 200)])
def test_odd_shapes_and_usage_block(cuda, oracle, shape, mining_name):
    """Reference batch sizes (usage/def.prototxt: 120 and 30 per rank), ragged tiles, the usage-block mining."""
    from gpu_harness import check_parity
    Q, world, D = shape
    mining = {"usage": synth.USAGE_MINING, "default": synth.DEFAULT_MINING,
              "local_rel": dict(synth.DEFAULT_MINING, ap_method=3, an_method=3, identsn=0.0, diffsn=-0.3, margin_diff=-0.01)}[mining_name]
    x, lab = synth.make_inputs(Q * world, D, seed=Q + world + D, imgs_per_class=2)
    check_parity(oracle, x, lab, Q, world, mining, capi.PREC_FP32_FP16X2, capi.GEMM_TCGEN05, tag=f"{shape} {mining_name}")


@pytest.mark.parametrize("flags", [0, capi.FLAG_LSEL_WARP])
@pytest.mark.parametrize("shape,ipc,sn", [((701, 1, 96), 3, (-0.3, -0.3)), ((512, 2, 64), 40, (-0.6, -0.05)), ((300, 1, 200), 150, (-0.2, -0.97)),
                                          ((1024, 1, 128), 4, (1.0, 5.0)), ((258, 1, 64), 129, (-0.5, -0.5))])
def test_local_relative_select_kernels(cuda, oracle, shape, ipc, sn, flags):
    """LOCAL RELATIVE_* on both sides through both per-row select kernels (block-per-row with the row in registers: rows of up to
    8192 columns; warp-per-row: npair_config.flags & NPAIR_FLAG_LSEL_WARP, and longer rows): few and many images per class (same-label
    lists beyond the 128-entry fast path), fractional and absolute positions, ragged row lengths.  Thresholds are bit-exact (level 2 of
    check_parity compares the selection-dependent results with the oracle run on the GPU's own similarities)."""
    from gpu_harness import check_parity
    Q, world, D = shape
    x, lab = synth.make_inputs(Q * world, D, seed=Q + D + ipc, imgs_per_class=ipc, noise=1.5)
    for apM, anM in ((capi.RELATIVE_HARD, capi.RELATIVE_HARD), (capi.RELATIVE_EASY, capi.RELATIVE_EASY)):
        mining = dict(margin_ident=0.01, margin_diff=-0.02, identsn=sn[0], diffsn=sn[1], ap_region=synth.LOCAL, ap_method=apM,
                      an_region=synth.LOCAL, an_method=anM)
        check_parity(oracle, x, lab, Q, world, mining, capi.PREC_FP32_FP16X2, capi.GEMM_TCGEN05, tag=f"lsel {shape} ipc{ipc} sn{sn} f{flags} m{apM}",
                     flags=flags)


@pytest.mark.parametrize("flags", [0, capi.FLAG_LSEL_WARP])
def test_local_relative_select_crowded_bins(cuda, oracle, flags):
    """Rows whose similarities crowd into one value bin: 300 exact copies of one embedding spread over different classes (equal keys)
    and 400 copies with 1e-6 noise (distinct keys, one bin): the block-per-row kernel's key-digit refinement; also one row of all-equal
    similarities apart from the copies (D = 1 direction)."""
    from gpu_harness import check_parity
    B, D = 1024, 64
    x, lab = synth.make_inputs(B, D, seed=4242, imgs_per_class=2, noise=1.5)
    rng = np.random.default_rng(7)
    idx = rng.permutation(B)
    x[idx[:300]] = x[idx[0]]
    x[idx[300:700]] = x[idx[300]] + 1e-6 * rng.standard_normal((400, D)).astype(np.float32)
    for sn in ((-0.3, -0.3), (-0.5, -0.62), (0.0, 40.0)):
        mining = dict(margin_ident=0.0, margin_diff=0.0, identsn=sn[0], diffsn=sn[1], ap_region=synth.LOCAL, ap_method=capi.RELATIVE_HARD,
                      an_region=synth.LOCAL, an_method=capi.RELATIVE_HARD)
        check_parity(oracle, x, lab, B, 1, mining, capi.PREC_FP32_FP16X2, capi.GEMM_TCGEN05, tag=f"crowded sn{sn} f{flags}", flags=flags)


def test_num_tops_layout(cuda, oracle):
    from gpu_harness import gpu_step_world
    Q, D = 64, 32
    x, lab = synth.make_inputs(Q, D, seed=4)
    t5 = gpu_step_world(x, lab, Q, 1, synth.DEFAULT_MINING, capi.PREC_FP32_FP16X2, capi.GEMM_TCGEN05, want_grad=False)["tops"][0]
    for nt in (1, 2, 3, 4):
        t = gpu_step_world(x, lab, Q, 1, synth.DEFAULT_MINING, capi.PREC_FP32_FP16X2, capi.GEMM_TCGEN05, num_tops=nt, want_grad=False)["tops"][0]
        exp = np.zeros(5, dtype=np.float32)
        exp[0] = t5[0]
        for k in range(1, nt - 1):
            exp[k] = t5[k]
        exp[nt - 1] = t5[4]          # last top is always the asum (overwrites the loss when nt == 1)
        np.testing.assert_allclose(t, exp, rtol=1e-6)


def test_error_codes(cuda):
    import torch
    Q, D = 16, 8
    x, lab = synth.make_inputs(Q, D, seed=1)
    xt, lt = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
    ctx = capi.Context(capi.make_config(Q, D, ap_method=capi.RELATIVE_HARD))      # identsn=-1 -> pos=-1
    with pytest.raises(capi.NpairError) as e:
        ctx.forward(xt, lt)
    assert e.value.code == -5
    ctx.close()
    ctx = capi.Context(capi.make_config(Q, D, an_region=capi.GLOBAL, an_method=capi.HARD))
    with pytest.raises(capi.NpairError) as e:
        ctx.forward(xt, torch.arange(Q, dtype=torch.float32).cuda())                 # no positive pair
    assert e.value.code == -4
    with pytest.raises(capi.NpairError) as e:
        ctx.backward(1.0, torch.empty_like(xt))                                      # no successful forward
    assert e.value.code == -6
    ctx.close()


def test_medium_headline_mining_fp32_faithful(cuda, oracle):
    """B=2048, D=512 with the usage-block mining in both fp32-faithful modes."""
    from gpu_harness import check_parity
    x, lab = synth.make_inputs(2048, 512, seed=20171225 + 5)
    for prec in (capi.PREC_FP32_FP16X2, capi.PREC_FP32_BF16X3):
        r = check_parity(oracle, x, lab, 2048, 1, synth.USAGE_MINING, prec, capi.GEMM_TCGEN05, tag=f"B2048 {PREC_NAME[prec]}")
        print("B2048", PREC_NAME[prec], r)


@pytest.mark.parametrize("B,D,prec", [(1024, 256, capi.PREC_FP32_FP16X2), (2048, 512, capi.PREC_BF16), (640, 200, capi.PREC_FP32_BF16X3)])
def test_pair_kernels_match_single_cta_bitwise(cuda, B, D, prec):
    """The CTA-pair (tcgen05 cta_group::2) similarity / gradient kernels against the single-CTA kernels (npair_config.flags): S, the
    fused row statistics, the tops and the gradient agree bit for bit."""
    torch = cuda
    x, lab = synth.make_inputs(B, D, 20171230, noise=2.5)
    dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
    res = []
    for flags in (0, capi.FLAG_SIM_1CTA | capi.FLAG_GRAD_1CTA):
        ctx = capi.Context(capi.make_config(B, D, sim_precision=prec, flags=flags, **synth.USAGE_MINING))
        dg = torch.full_like(dx, float("nan"))
        tops = ctx.forward(dx, dl)
        ctx.backward(1.0, dg)
        torch.cuda.synchronize()
        res.append((np.array(tops, np.float32), ctx.debug_read(0, B * B), np.stack([ctx.debug_read(w, B) for w in (3, 4, 5, 8, 9)]), dg.cpu().numpy()))
        ctx.close()
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_mma_symmetry_self_check(cuda):
    """npair_create's one-off device check behind the row-record exchange: a similarity matrix computed with every tile is bitwise
    symmetric in each operand format (if it ever is not, world > 1 contexts use the reduce-scatter form)."""
    for prec in PRECS:
        assert capi.lib().npair_debug_mma_symmetric(prec) == 1, PREC_NAME[prec]

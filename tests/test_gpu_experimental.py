"""Opt-in GPU checks of kernels that are compiled in but not on the default path yet (NPAIR_RUN_EXPERIMENTAL=1).
They compare the opt-in path with the default path of the same build on the same inputs."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("NPAIR_RUN_EXPERIMENTAL") != "1",
                                                  reason="experimental kernels: set NPAIR_RUN_EXPERIMENTAL=1")]


def _run(B, D, mining, env):
    import torch
    from npairloss_b200 import capi, synth
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        x, lab = synth.make_inputs(B, D, 1234 + B, noise=2.5)
        ctx = capi.Context(capi.make_config(B, D, **mining))       # the switches are read at npair_create
        dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
        dg = torch.empty_like(dx)
        tops = ctx.forward(dx, dl)
        ctx.backward(1.0, dg)
        A, T = ctx.debug_read(6, B), ctx.debug_read(7, B)
        tops2 = ctx.forward(dx, dl)                                 # second step: counters were reset
        return np.array(tops), np.array(tops2), A, T, dg.cpu().numpy()
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v


@pytest.mark.parametrize("B,D", [(256, 64), (1000, 200), (2048, 128), (8192, 512)])
@pytest.mark.parametrize("mining", ["usage", "default", "hard"])
@pytest.mark.parametrize("tiles_mode", ["1", "2"])          # 1: symmetric walk (upper triangle), 2: plain tile walk
def test_tile_row_pass_matches_row_pass(B, D, mining, tiles_mode):
    from npairloss_b200 import synth
    m = {"usage": synth.USAGE_MINING, "default": synth.DEFAULT_MINING,
         "hard": dict(synth.DEFAULT_MINING, ap_method=synth.HARD, an_method=synth.HARD)}[mining]
    t0, t0b, A0, T0, g0 = _run(B, D, m, {"NPAIR_LSE_TILES": "0"})
    t1, t1b, A1, T1, g1 = _run(B, D, m, {"NPAIR_LSE_TILES": tiles_mode})
    np.testing.assert_allclose(t1, t0, rtol=2e-6, atol=1e-7)        # other summation order of T only
    np.testing.assert_allclose(t1b, t1, rtol=0, atol=0)             # deterministic, state reset between steps
    np.testing.assert_allclose(A1, A0, rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(T1, T0, rtol=2e-6, atol=1e-30)
    assert np.linalg.norm(g1 - g0) <= 2e-6 * max(np.linalg.norm(g0), 1e-30)


def _grad_in_subprocess(B, D, prec, env):
    """Gradient of one forward+backward in a fresh process (the experimental switches are read once per process / context)."""
    import subprocess, sys, tempfile
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from npairloss_b200 import capi, synth
B, D, prec = %d, %d, %d
x, lab = synth.make_inputs(B, D, 99, noise=2.5)
ctx = capi.Context(capi.make_config(B, D, sim_precision=prec, **synth.USAGE_MINING))
dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
dg = torch.empty_like(dx)
for _ in range(3):                       # several launches: epochs / counters must carry over correctly
    ctx.forward(dx, dl); ctx.backward(1.0, dg)
np.save(sys.argv[1], dg.cpu().numpy())
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), B, D, prec)
    with tempfile.NamedTemporaryFile(suffix=".npy", delete=False) as f:
        path = f.name
    subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, **env), check=True, timeout=300)
    g = np.load(path); os.unlink(path)
    return g


@pytest.mark.parametrize("B,D,prec", [(512, 64, 2), (1000, 200, 2), (2048, 1024, 0), (8192, 512, 2), (8192, 512, 1)])
@pytest.mark.parametrize("one_ex2", ["0", "1"])
def test_stream_k_gradient_matches(B, D, prec, one_ex2):
    """NPAIR_GRAD_STREAMK=1 (grad_streamk.cuh) against the default CTA-pair kernel; only the split of the K sum differs."""
    g0 = _grad_in_subprocess(B, D, prec, {"NPAIR_GRAD_STREAMK": "0", "NPAIR_GRAD_ONE_EX2": "0"})
    g1 = _grad_in_subprocess(B, D, prec, {"NPAIR_GRAD_STREAMK": "1", "NPAIR_GRAD_ONE_EX2": one_ex2})
    tol = (2e-6 if one_ex2 == "0" else 5e-6) if prec != 1 else 2e-3
    assert np.linalg.norm(g1 - g0) <= tol * max(np.linalg.norm(g0), 1e-30)


@pytest.mark.parametrize("B,D,prec", [(1000, 200, 2), (2048, 512, 0), (8192, 512, 2), (8192, 512, 1)])
def test_single_exponential_producer_matches(B, D, prec):
    """NPAIR_GRAD_ONE_EX2=1: e2 formed from e1 and per-row / per-column constants (grad_fused.cuh).  NOTE the switch is read once
    per process: run this test in its own pytest process, once with the variable set and once without, or rely on the two
    sub-processes below."""
    import subprocess, sys, json, tempfile
    from npairloss_b200 import synth
    code = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %r)
from npairloss_b200 import capi, synth
B, D, prec = %d, %d, %d
x, lab = synth.make_inputs(B, D, 99, noise=2.5)
ctx = capi.Context(capi.make_config(B, D, sim_precision=prec, **synth.USAGE_MINING))
dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
dg = torch.empty_like(dx)
ctx.forward(dx, dl); ctx.backward(1.0, dg)
np.save(sys.argv[1], dg.cpu().numpy())
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), B, D, prec)
    outs = []
    for flag in ("0", "1"):
        with tempfile.NamedTemporaryFile(suffix=".npy", delete=False) as f:
            path = f.name
        env = dict(os.environ, NPAIR_GRAD_ONE_EX2=flag)
        subprocess.run([sys.executable, "-c", code, path], env=env, check=True, timeout=300)
        outs.append(np.load(path)); os.unlink(path)
    g0, g1 = outs
    tol = 5e-6 if prec != 1 else 2e-3
    assert np.linalg.norm(g1 - g0) <= tol * max(np.linalg.norm(g0), 1e-30)


def test_torch_api_matches_capi():
    """npairloss_b200.torch_api (autograd surface) against direct C-ABI calls."""
    import torch
    from npairloss_b200 import capi, synth, torch_api
    B, D = 512, 128
    x, lab = synth.make_inputs(B, D, 7, noise=2.5)
    ctx = capi.Context(capi.make_config(B, D, **synth.USAGE_MINING))
    dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
    dg = torch.empty_like(dx)
    tops = ctx.forward(dx, dl); ctx.backward(0.5, dg)
    m = torch_api.NPairLoss(**synth.USAGE_MINING)
    xr = dx.clone().requires_grad_(True)
    loss, t = m(xr, dl)
    (0.5 * loss).backward()
    np.testing.assert_allclose(t.cpu().numpy(), np.array(tops, np.float32), rtol=0, atol=0)
    np.testing.assert_allclose(xr.grad.cpu().numpy(), dg.cpu().numpy(), rtol=0, atol=0)


def test_full_size_properties_headline():
    """BASELINE.json's full size (B=8192, D=512, usage-block mining) is beyond the oracle's reach in a test; the domain's
    size-independent properties stand in (promote to test_gpu_parity.py once it has run on a B200):
    S bitwise symmetric, sample-permutation equivariance, label renaming invariance, gradient linear in the loss weight."""
    import torch
    from npairloss_b200 import capi, synth
    B, D = 8192, 512
    x, lab = synth.make_inputs(B, D, 20171225 + 5, noise=2.5)
    ctx = capi.Context(capi.make_config(B, D, **synth.USAGE_MINING))

    def step(xx, ll, lw=1.0):
        dx, dl = torch.from_numpy(np.ascontiguousarray(xx)).cuda(), torch.from_numpy(np.ascontiguousarray(ll)).cuda()
        dg = torch.empty_like(dx)
        tops = ctx.forward(dx, dl); ctx.backward(lw, dg)
        return np.array(tops, np.float32), dg.cpu().numpy()

    t0, g0 = step(x, lab)
    S = ctx.debug_read(0, B * B).reshape(B, B)
    assert np.array_equal(S, S.T)
    perm = np.random.default_rng(5).permutation(B)
    t1, g1 = step(x[perm], lab[perm])
    np.testing.assert_allclose(t1, t0, rtol=1e-5, atol=1e-7)
    assert np.linalg.norm(g1 - g0[perm]) <= 1e-5 * np.linalg.norm(g0)
    t2, g2 = step(x, lab * 3.0 + 17.0)
    np.testing.assert_array_equal(t2, t0); np.testing.assert_array_equal(g2, g0)
    t3, g3 = step(x, lab, lw=-0.5)
    np.testing.assert_array_equal(t3, t0)
    assert np.linalg.norm(g3 + 0.5 * g0) <= 1e-6 * np.linalg.norm(g0)


def test_forward_backward_single_sync_matches_two_calls():
    """npair_forward_backward against npair_forward + npair_backward: bit-identical tops and gradient, error codes preserved."""
    import torch
    from npairloss_b200 import capi, synth
    for B, D in ((512, 128), (2048, 512)):
        x, lab = synth.make_inputs(B, D, 5, noise=2.5)
        ctx = capi.Context(capi.make_config(B, D, **synth.USAGE_MINING))
        dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
        g0, g1 = torch.empty_like(dx), torch.empty_like(dx)
        t0 = ctx.forward(dx, dl); ctx.backward(0.7, g0)
        t1 = ctx.forward_backward(dx, dl, 0.7, g1)
        assert t0 == t1 and torch.equal(g0, g1)
        t2 = ctx.forward(dx, dl); ctx.backward(0.7, g0)             # the two-call path still works afterwards
        assert t2 == t0
        ctx.close()
    ctx = capi.Context(capi.make_config(16, 8, ap_method=capi.RELATIVE_HARD))   # identsn = -1 -> pos out of range
    x, lab = synth.make_inputs(16, 8, 1)
    dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
    with pytest.raises(capi.NpairError) as e:
        ctx.forward_backward(dx, dl, 1.0, torch.empty_like(dx))
    assert e.value.code == -5

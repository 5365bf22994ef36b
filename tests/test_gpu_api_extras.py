"""GPU checks of the API surfaces around the two core calls: the PyTorch autograd wrapper, npair_forward_backward
(one host synchronisation), and the size-independent properties of the path at BASELINE.json's full size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


def test_torch_true_gradient_is_the_derivative_of_the_loss():
    """true_gradient=True: central differences of the module's own loss (RAND mining: every pair selected, the loss is smooth)."""
    import torch
    from npairloss_b200 import synth, torch_api
    B, D = 32, 16
    x, lab = synth.make_inputs(B, D, 3, noise=1.0)
    m = torch_api.NPairLoss(true_gradient=True, sim_precision=0, **synth.DEFAULT_MINING)      # bf16x3: no pre-scale step to perturb
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    lt = torch.from_numpy(lab).cuda()
    loss, _ = m(xt, lt)
    loss.backward()
    g = xt.grad.cpu().numpy()
    rng = np.random.default_rng(0)
    for _ in range(6):
        i, j = int(rng.integers(B)), int(rng.integers(D))
        h = 2e-2
        xp, xm = x.copy(), x.copy()
        xp[i, j] += h; xm[i, j] -= h
        lp = m(torch.from_numpy(xp).cuda(), lt)[0].item()
        lm = m(torch.from_numpy(xm).cuda(), lt)[0].item()
        fd = (lp - lm) / (2 * h)
        assert abs(fd - g[i, j]) <= 2e-3 * max(abs(fd), 1e-2), (i, j, fd, g[i, j])


def test_full_size_properties_headline():
    """BASELINE.json's full size (B=8192, D=512, usage-block mining; the oracle comparison at this size is
    tests/test_gpu_baseline_configs.py): the domain's size-independent properties -- S bitwise symmetric, sample-permutation
    equivariance, label renaming invariance, gradient linear in the loss weight."""
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
    # a permutation only reorders the fp32 sums over the sample index: two results that are each within 1e-5 of the exact value
    assert np.linalg.norm(g1 - g0[perm]) <= 2e-5 * np.linalg.norm(g0)
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

"""CPU tests that pin the oracle (parity is unpinned by the reference, so the oracle is pinned by
 (i) the hand KAT of SURVEY 9.3, (ii) finite differences, (iii) an independent NumPy restatement,
 (iv) multi-rank identities)."""
import itertools
import math

import numpy as np
import pytest

from npairloss_b200 import synth
from oracle import npair_oracle_np as onp

METHODS = [0, 1, 2, 3, 4]
REGIONS = [0, 1]


def test_kat_survey_9_3(oracle):
    # k=1, Q=4, D=2, labels [0,0,1,1], x0=x1=(1,0), x2=x3=(0,1), defaults, lw=1
    x = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=np.float32)
    lab = np.array([0, 0, 1, 1], dtype=np.float32)
    cfg = oracle.make_config(4, 2)
    tops, dx = oracle.step_world(x, lab, cfg, 1.0)
    assert tops[0, 0] == pytest.approx(math.log(1 + 2 / math.e), rel=1e-6)       # 0.5514447139
    assert tops[0, 1] == 1.0 and tops[0, 2] == 1.0 and tops[0, 3] == 1.0
    assert tops[0, 4] == pytest.approx(1.0, rel=1e-7)
    g = 0.1059707788
    exp = np.array([[-g, g], [-g, g], [g, -g], [g, -g]], dtype=np.float32)
    np.testing.assert_allclose(dx, exp, rtol=2e-6)
    # the NumPy restatement gives the same
    t2, d2 = onp.step_world(x, lab, 4, 1)
    np.testing.assert_allclose(t2[0], tops[0], rtol=1e-6)
    np.testing.assert_allclose(d2, exp, rtol=2e-6)


def test_pos_fp32_semantics(oracle):
    # fp32 index arithmetic
    assert oracle.pos(-0.3, 67092480) == 46964736          # exact arithmetic would give 46964735
    assert oracle.pos(-0.0, 10) == 9                       # -0.0 >= 0 -> size-1
    assert oracle.pos(0.0, 10) == 9
    assert oracle.pos(2.0, 10) == 7
    assert oracle.pos(-1.0, 10) == -1                      # proto default -> out of range (UB upstream)
    assert oracle.pos(20.0, 10) < 0
    for sn, size in [(-0.3, 100), (-0.5, 7), (-0.99, 1000), (-0.25, 8191)]:
        assert oracle.pos(sn, size) == onp.pos_index(sn, size)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_cpp_matches_numpy_all_modes(oracle, world):
    """Two independent restatements agree for every (region, method) combination."""
    Q, D = 24, 16
    N = Q * world
    x, lab = synth.make_inputs(N, D, seed=7 + world, imgs_per_class=3, noise=0.7)
    n = 0
    for apR, apM, anR, anM in itertools.product(REGIONS, METHODS, REGIONS, METHODS):
        kw = dict(margin_ident=0.02, margin_diff=-0.03, identsn=-0.4, diffsn=-0.3,
                  ap_region=apR, ap_method=apM, an_region=anR, an_method=anM)
        cfg = oracle.make_config(Q, D, world=world, **kw)
        tops, dx = oracle.step_world(x, lab, cfg, 0.7)
        t2, d2 = onp.step_world(x, lab, Q, world, 0.7, **kw)
        np.testing.assert_allclose(tops, t2, rtol=2e-6, atol=1e-7, err_msg=str(kw))
        np.testing.assert_allclose(dx, d2, rtol=1e-4, atol=2e-7, err_msg=str(kw))
        assert np.linalg.norm(dx - d2) <= 2e-6 * max(np.linalg.norm(d2), 1e-12), kw
        n += 1
    assert n == 100


def test_faithful_and_fast_sorts_agree(oracle):
    Q, D, world = 40, 32, 2
    x, lab = synth.make_inputs(Q * world, D, seed=3, imgs_per_class=4)
    for apR, apM, anR, anM in [(0, 3, 1, 0), (1, 3, 1, 3), (0, 4, 0, 4), (0, 0, 0, 1), (1, 2, 1, 2)]:
        kw = dict(margin_diff=-0.05, identsn=-0.2, diffsn=-0.3, ap_region=apR, ap_method=apM, an_region=anR, an_method=anM)
        a = oracle.step_world(x, lab, oracle.make_config(Q, D, world=world, faithful_sorts=1, **kw))
        b = oracle.step_world(x, lab, oracle.make_config(Q, D, world=world, faithful_sorts=0, num_threads=2, **kw))
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
        c = oracle.step_world(x, lab, oracle.make_config(Q, D, world=world, accum_double=0, **kw))
        np.testing.assert_allclose(a[0], c[0], rtol=1e-5, atol=1e-6)
        assert np.linalg.norm(a[1] - c[1]) <= 1e-5 * np.linalg.norm(a[1])


def test_finite_difference_half_gradient(oracle):
    """dX == 1/2 * dLoss/dX for k=1 where selection is locally constant (RAND/RAND): SURVEY Q8."""
    Q, D = 12, 6
    x, lab = synth.make_inputs(Q, D, seed=11, imgs_per_class=2)
    x = x.astype(np.float64)

    def loss64(xx):
        S = xx @ xx.T
        eq = lab[:, None] == lab[None, :]
        ns = ~np.eye(Q, dtype=bool)
        mx = np.where(ns, S, -np.inf).max(axis=1, keepdims=True)
        E = np.exp(S - mx)
        A = np.where(eq & ns, E, 0).sum(axis=1)
        B = np.where(~eq & ns, E, 0).sum(axis=1)
        return -np.mean(np.log(A / (A + B)))

    cfg = oracle.make_config(Q, D)
    tops, dx = oracle.step_world(x.astype(np.float32), lab, cfg, 1.0)
    assert tops[0, 0] == pytest.approx(loss64(x), rel=2e-6)
    g = np.zeros_like(x)
    h = 1e-6
    for i in range(Q):
        for d in range(D):
            xp = x.copy(); xp[i, d] += h
            xm = x.copy(); xm[i, d] -= h
            g[i, d] = (loss64(xp) - loss64(xm)) / (2 * h)
    np.testing.assert_allclose(dx, 0.5 * g, rtol=2e-4, atol=2e-7)


def test_tops_layout_quirks(oracle):
    """Q10: last top is always the feature asum; a single top is overwritten by it."""
    Q, D = 16, 8
    x, lab = synth.make_inputs(Q, D, seed=5)
    t5, _ = oracle.forward(x, lab, oracle.make_config(Q, D, num_tops=5))
    t3, _ = oracle.forward(x, lab, oracle.make_config(Q, D, num_tops=3))
    t2, _ = oracle.forward(x, lab, oracle.make_config(Q, D, num_tops=2))
    t1, _ = oracle.forward(x, lab, oracle.make_config(Q, D, num_tops=1))
    asum = t5[4]
    assert asum == pytest.approx(np.abs(x).sum() / Q, rel=1e-6)
    assert t3[0] == t5[0] and t3[1] == t5[1] and t3[2] == asum
    assert t2[0] == t5[0] and t2[1] == asum
    assert t1[0] == asum


def test_error_codes_for_reference_ub(oracle):
    Q, D = 8, 4
    x, lab = synth.make_inputs(Q, D, seed=1)
    # proto-default identsn=-1 with a relative method -> pos = -1 (UB upstream) -> error 3
    with pytest.raises(oracle.OracleError) as e:
        oracle.forward(x, lab, oracle.make_config(Q, D, ap_method=3))
    assert e.value.code == 3
    # all labels distinct -> no positive pair -> GLOBAL HARD AN indexes an empty list -> error 2
    with pytest.raises(oracle.OracleError) as e:
        oracle.forward(x, np.arange(Q, dtype=np.float32), oracle.make_config(Q, D, an_region=0, an_method=0))
    assert e.value.code == 2


def test_relative_negative_order_statistic_clamps(oracle):
    """Q4: a negative order statistic becomes -FLT_MAX."""
    Q, D = 16, 8
    x, lab = synth.make_inputs(Q, D, seed=9)
    # diffsn=-0.9 -> a low (negative) negative-pair similarity -> clamp -> RELATIVE_HARD AN selects all negatives
    _, st = oracle.forward(x, lab, oracle.make_config(Q, D, an_method=3, diffsn=-0.9))
    assert np.all(st["nega_thr"] == -np.finfo(np.float32).max)
    assert np.all(st["diff_num"] == Q - 2)


def test_multirank_loss_is_per_rank_and_grad_blend(oracle):
    """Q8/Q9: with k ranks the transposed term is divided by k; emulate and compare to the closed form."""
    Q, D, k = 10, 8, 2
    N = Q * k
    x, lab = synth.make_inputs(N, D, seed=21)
    cfg = oracle.make_config(Q, D, world=k)
    tops, dx = oracle.step_world(x, lab, cfg, 1.0)
    G = np.zeros((N, N))
    for r in range(k):
        _, st = onp.forward(x, lab, Q, k, r)
        G[r * Q:(r + 1) * Q] = onp.grad_weights(st, Q, 1.0)
    ref = 0.5 * G @ x.astype(np.float64) + 0.5 / k * G.T @ x.astype(np.float64)
    assert np.linalg.norm(dx - ref) <= 2e-6 * np.linalg.norm(ref)


def test_l2normalize_statement(oracle):
    """The L2Normalize producer (usage/def.prototxt:115-120; source not in the reference tree): unit rows, zero rows kept,
    backward = the Jacobian of x / ||x|| (finite differences in fp64)."""
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((6, 9)) * rng.uniform(0.1, 50, size=(6, 1))).astype(np.float32)
    x[4] = 0
    y, inv = oracle.l2normalize_forward(x)
    np.testing.assert_allclose(np.linalg.norm(y[[0, 1, 2, 3, 5]].astype(np.float64), axis=1), 1.0, rtol=1e-6)
    assert not y[4].any() and inv[4] == 0
    dy = rng.standard_normal((6, 9)).astype(np.float32)
    dx = oracle.l2normalize_backward(y, inv, dy)

    def f(xx):
        n = np.linalg.norm(xx, axis=1, keepdims=True)
        return float((np.where(n > 0, xx / np.where(n > 0, n, 1), 0) * dy).sum())
    x64 = x.astype(np.float64)
    for i in (0, 1, 2, 3, 5):
        for j in range(9):
            h = 1e-6 * max(1.0, abs(x64[i, j]))
            e = np.zeros_like(x64); e[i, j] = h
            fd = (f(x64 + e) - f(x64 - e)) / (2 * h)
            assert abs(fd - dx[i, j]) <= 1e-5 * max(1e-3, abs(fd)) + 1e-7, (i, j, fd, dx[i, j])
    assert not dx[4].any()

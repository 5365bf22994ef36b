"""Size-independent properties of the oracle (the domain's invariances), on top of test_oracle.py's pins:
   sample-permutation equivariance, label-renaming invariance, linearity in the loss weight, rank-count identities of LOCAL
   mining.  These are the properties the GPU parity tests re-use at sizes the oracle cannot reach."""
import numpy as np
import pytest

from npairloss_b200 import synth

MININGS = {
    "usage": synth.USAGE_MINING,
    "rand": synth.DEFAULT_MINING,
    "hard_hard": dict(synth.DEFAULT_MINING, ap_method=synth.HARD, an_method=synth.HARD),
    "local_relative": dict(synth.USAGE_MINING, ap_region=synth.LOCAL, an_region=synth.LOCAL, ap_method=synth.RELATIVE_HARD,
                           an_method=synth.RELATIVE_EASY, identsn=-0.4, diffsn=-0.25, margin_diff=0.0),
    "global_easy": dict(synth.DEFAULT_MINING, ap_region=synth.GLOBAL, an_region=synth.GLOBAL, ap_method=synth.EASY, an_method=synth.EASY),
}


def _inputs(B=48, D=16, seed=3, imgs=3):
    x, lab = synth.make_inputs(B, D, seed, imgs_per_class=imgs, noise=2.0)
    return x, lab


@pytest.mark.parametrize("mining", sorted(MININGS))
def test_permutation_equivariance(oracle, mining):
    """Re-ordering the samples (features and labels together) permutes the gradient rows and leaves every top unchanged:
    nothing in .cu:207-499 depends on the position of a sample except the self-pair exclusion, which moves with it."""
    x, lab = _inputs()
    cfg = oracle.make_config(len(lab), x.shape[1], **MININGS[mining])
    t0, g0 = oracle.step_world(x, lab, cfg, 1.0)
    perm = np.random.default_rng(11).permutation(len(lab))
    t1, g1 = oracle.step_world(x[perm], lab[perm], cfg, 1.0)
    np.testing.assert_allclose(t1, t0, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(g1, g0[perm], rtol=2e-5, atol=2e-8)


@pytest.mark.parametrize("mining", sorted(MININGS))
def test_label_renaming_invariance(oracle, mining):
    """Labels only enter through equality (GetLabelDiffMtx, .cu:44-66): any injective renaming gives identical outputs."""
    x, lab = _inputs()
    cfg = oracle.make_config(len(lab), x.shape[1], **MININGS[mining])
    t0, g0 = oracle.step_world(x, lab, cfg, 1.0)
    t1, g1 = oracle.step_world(x, (lab * 7.0 + 1000.0).astype(np.float32), cfg, 1.0)
    np.testing.assert_array_equal(t1, t0)
    np.testing.assert_array_equal(g1, g0)


@pytest.mark.parametrize("lw", [0.0, 0.25, -3.0])
def test_gradient_is_linear_in_loss_weight(oracle, lw):
    """G = (lw/Q)(-W1+W2+W3) (.cu:435-460): the gradient scales with top[0]'s diff, the forward does not see it."""
    x, lab = _inputs()
    cfg = oracle.make_config(len(lab), x.shape[1], **MININGS["usage"])
    t1, g1 = oracle.step_world(x, lab, cfg, 1.0)
    tl, gl = oracle.step_world(x, lab, cfg, lw)
    np.testing.assert_array_equal(tl, t1)
    np.testing.assert_allclose(gl, np.float32(lw) * g1, rtol=1e-5, atol=1e-9)   # lw enters before the fp32 contractions


def test_feature_scaling_moves_only_the_asum_for_rand_mining(oracle):
    """With RAND/RAND (every pair selected, no thresholds) scaling all features by a power of two scales S by its square;
    asum (.cu:400) scales exactly and the retrieval tops (rank statistics of S) do not move."""
    x, lab = _inputs()
    cfg = oracle.make_config(len(lab), x.shape[1], **MININGS["rand"])
    t0, _ = oracle.step_world(x, lab, cfg, 1.0)
    t1, _ = oracle.step_world(x * np.float32(0.5), lab, cfg, 1.0)
    np.testing.assert_array_equal(t1[:, 1:4], t0[:, 1:4])
    np.testing.assert_allclose(t1[:, 4], 0.5 * t0[:, 4], rtol=1e-6)


@pytest.mark.parametrize("world", [2, 4])
def test_local_mining_rows_do_not_depend_on_the_sharding(oracle, world):
    """LOCAL thresholds are per anchor row, so the ROW part of the gradient (local_diff, .cu:448-453) and the per-row loss terms
    are the same however the anchors are sharded; what changes with k is the per-rank normaliser Q (.cu:385, :427) and the
    1/k blend (.cu:492-497).  Checked through the loss: mean over ranks of the per-rank losses == the k=1 loss."""
    x, lab = _inputs(B=48)
    m = MININGS["hard_hard"]
    c1 = oracle.make_config(48, x.shape[1], **m)
    ck = oracle.make_config(48 // world, x.shape[1], world=world, **m)
    t1, _ = oracle.step_world(x, lab, c1, 1.0)
    tk, _ = oracle.step_world(x, lab, ck, 1.0)
    assert tk.shape == (world, 5)
    np.testing.assert_allclose(tk[:, 0].mean(), t1[0, 0], rtol=2e-6)
    np.testing.assert_allclose(tk[:, 1:4].mean(axis=0), t1[0, 1:4], rtol=1e-6)

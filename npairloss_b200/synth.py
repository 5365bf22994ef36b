"""Synthetic inputs of SURVEY.md section 8(d): B/imgs classes x imgs images, label[g] = g // imgs (pairs
contiguous like MultibatchData, usage/def.prototxt:21-27), x = normalize(c_label + noise*eps)."""
from __future__ import annotations

import numpy as np

# (ap_region, ap_method, an_region, an_method, margin_ident, margin_diff, identsn, diffsn)
GLOBAL, LOCAL = 0, 1
HARD, EASY, RAND, RELATIVE_HARD, RELATIVE_EASY = 0, 1, 2, 3, 4

# usage/def.prototxt:137-146 -- the reference's own layer block
USAGE_MINING = dict(margin_ident=0.0, margin_diff=-0.05, identsn=-0.0, diffsn=-0.3,
                    ap_region=GLOBAL, ap_method=RELATIVE_HARD, an_region=LOCAL, an_method=HARD)
DEFAULT_MINING = dict(margin_ident=0.0, margin_diff=0.0, identsn=-1.0, diffsn=-1.0,
                      ap_region=LOCAL, ap_method=RAND, an_region=LOCAL, an_method=RAND)

# BASELINE.json configs (index = config_index of the seed rule)
# noise: SURVEY 8(d) proposes 1.0 (positive cos ~0.5, negatives ~N(0,1/sqrt(D))).  With a margin-based HARD negative
# rule that leaves NO negative selected at D >= 256 (max negative ~0.2 < 0.5 - 0.05): loss and gradient are exactly
# zero, a degenerate benchmark.  The mining configs therefore use noise 2.5 (positive cos ~ 1/(1+2.5^2) = 0.14), which
# selects a few percent of the negatives; the RAND/RAND configs keep 1.0.  Kernel cost does not depend on this choice.
CONFIGS = {
    "C1": dict(B=64, D=128, world=1, mining=DEFAULT_MINING, idx=0, noise=1.0),
    "C2": dict(B=512, D=128, world=1, mining=DEFAULT_MINING, idx=1, noise=1.0),
    "C3": dict(B=4096, D=512, world=1, mining=dict(DEFAULT_MINING, an_method=HARD), idx=2, noise=2.5),
    "C4": dict(B=8192, D=1024, world=8, mining=dict(USAGE_MINING), idx=3, noise=2.5),
    "C5": dict(B=65536, D=256, world=8, mining=dict(DEFAULT_MINING, ap_method=HARD, an_method=HARD), idx=4, noise=2.5),
    "HL": dict(B=8192, D=512, world=1, mining=dict(USAGE_MINING), idx=5, noise=2.5),
}


def make_inputs(B: int, D: int, seed: int, imgs_per_class: int = 2, noise: float = 1.0):
    rng = np.random.default_rng(seed)
    n_cls = (B + imgs_per_class - 1) // imgs_per_class
    labels = (np.arange(B) // imgs_per_class).astype(np.float32)
    centers = rng.standard_normal((n_cls, D)).astype(np.float32)
    eps = rng.standard_normal((B, D)).astype(np.float32)
    x = centers[labels.astype(np.int64)] + np.float32(noise) * eps
    x /= np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32), labels


def config_inputs(name: str):
    c = CONFIGS[name]
    return make_inputs(c["B"], c["D"], 20171225 + c["idx"], noise=c["noise"])

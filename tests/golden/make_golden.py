"""Generates tests/golden/npair_golden.npz.

PARITY UNPINNED BY THE REFERENCE: quziyan/NPairLoss ships no tests or golden vectors and cannot be built here (private
Caffe fork + MPI), so these vectors come from the CPU oracle (oracle/npair_oracle.cpp, double accumulation, faithful
sorts), which is itself pinned by the hand KAT of SURVEY.md 9.3, finite differences and the independent NumPy
restatement (tests/test_oracle.py).  They freeze the oracle's behaviour (regression pin) and give the GPU tests a
fixture that does not need liboracle at run time.

    python tests/golden/make_golden.py        # rewrites npair_golden.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from npairloss_b200 import synth  # noqa: E402
from oracle import oracle_lib as o  # noqa: E402

CASES = [
    # name, Q, world, D, imgs/class, noise, loss_weight, mining
    ("kat_9_3", 4, 1, 2, 2, None, 1.0, synth.DEFAULT_MINING),
    ("default_q32", 32, 1, 16, 2, 1.0, 1.0, synth.DEFAULT_MINING),
    ("usage_q48", 48, 1, 24, 2, 2.5, 1.0, synth.USAGE_MINING),
    ("usage_q24_w2", 24, 2, 24, 2, 2.5, 0.5, synth.USAGE_MINING),
    ("local_rel_q40", 40, 1, 12, 4, 1.5, 1.0, dict(synth.DEFAULT_MINING, ap_method=3, an_method=3, identsn=1.0, diffsn=-0.3, margin_diff=-0.01)),
    ("global_rel_q36_w3", 12, 3, 20, 3, 1.5, 2.0, dict(synth.DEFAULT_MINING, ap_region=0, an_region=0, ap_method=4, an_method=3, identsn=-0.5, diffsn=-0.7, margin_ident=0.01)),
    ("hard_easy_q30", 30, 1, 10, 3, 1.5, 1.0, dict(synth.DEFAULT_MINING, ap_method=0, an_method=1, margin_ident=0.05, margin_diff=0.02)),
    ("global_hard_q16_w2", 16, 2, 8, 2, 2.0, 1.0, dict(synth.DEFAULT_MINING, ap_region=0, ap_method=1, an_region=0, an_method=0, margin_diff=-0.1)),
]


def main():
    out = {}
    names = []
    for name, Q, world, D, imgs, noise, lw, mining in CASES:
        N = Q * world
        if name == "kat_9_3":
            x = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=np.float32)
            lab = np.array([0, 0, 1, 1], dtype=np.float32)
        else:
            x, lab = synth.make_inputs(N, D, seed=abs(hash(name)) % 1000 + 11 if False else sum(map(ord, name)), imgs_per_class=imgs, noise=noise)
        cfg = o.make_config(Q, D, world=world, accum_double=1, faithful_sorts=1, **mining)
        tops, dx = o.step_world(x, lab, cfg, lw)
        names.append(name)
        out[f"{name}/x"] = x
        out[f"{name}/label"] = lab
        out[f"{name}/tops"] = tops
        out[f"{name}/dx"] = dx
        out[f"{name}/meta"] = np.array([Q, world, D, lw, mining["margin_ident"], mining["margin_diff"], mining["identsn"], mining["diffsn"],
                                        mining["ap_region"], mining["ap_method"], mining["an_region"], mining["an_method"]], dtype=np.float64)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "npair_golden.npz"), **out)
    print("wrote", len(names), "cases")


if __name__ == "__main__":
    main()

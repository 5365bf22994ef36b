"""Tuning helper (not a test): similarity-GEMM phase time only; tolerates debug builds whose statistics are incomplete.
   NPAIR_LIB=<variant.so> python tests/tune_sim.py [B] [D] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from npairloss_b200 import capi, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
prec = {"fp16x2": 2, "bf16": 1, "bf16x3": 0}[sys.argv[3] if len(sys.argv) > 3 else "fp16x2"]
x, lab = synth.make_inputs(B, D, 20171230, noise=2.5)
ctx = capi.Context(capi.make_config(B, D, sim_precision=prec, **synth.USAGE_MINING))
dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
ctx.profile_enable(True)
acc = []
for i in range(25):
    try:
        ctx.forward(dx, dl)
    except Exception as e:
        pass
    if i >= 5: acc.append(ctx.profile_read()[2])
print(os.path.basename(os.environ.get("NPAIR_LIB", "default")), "1cta" if os.environ.get("NPAIR_SIM_1CTA") == "1" else "pair", sys.argv[1:],
      "sim=%.1fus (min %.1f)" % (np.mean(acc) * 1e3, np.min(acc) * 1e3), flush=True)

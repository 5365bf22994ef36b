"""Tuning helper (not a test): per-phase CUDA-event times of one configuration.
   NPAIR_LIB=<variant.so> python tests/tune_phases.py [B] [D] [precision] [world-emulated=1]"""
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
dg = torch.empty_like(dx)
for _ in range(5):
    ctx.forward(dx, dl); ctx.backward(1.0, dg)
ctx.profile_enable(True)
acc = np.zeros(9); n = 20
for _ in range(n):
    ctx.forward(dx, dl); ctx.backward(1.0, dg); acc += np.array(ctx.profile_read())
acc /= n
names = ["allg", "prep", "sim", "thr", "row", "build", "grad", "gradT", "bwdx"]
print(os.environ.get("NPAIR_LIB", "default"), sys.argv[1:], " ".join(f"{k}={v*1e3:.1f}us" for k, v in zip(names, acc)), f"sum={acc.sum()*1e3:.1f}us")

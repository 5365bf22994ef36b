"""Tuning helper (not a test): per-phase CUDA-event times of one configuration.
   NPAIR_LIB=<variant.so> python tests/tune_phases.py [B] [D] [precision] [grad_chunk_cols (-1 = unchunked)] [flags] [mining usage|rand|relative|grel]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from npairloss_b200 import capi, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
prec = {"fp16x2": 2, "bf16": 1, "bf16x3": 0}[sys.argv[3] if len(sys.argv) > 3 else "fp16x2"]
x, lab = synth.make_inputs(B, D, 20171230, noise=2.5)
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
mname = sys.argv[6] if len(sys.argv) > 6 else "usage"
mining = {"usage": synth.USAGE_MINING, "rand": synth.DEFAULT_MINING,
          "relative": dict(synth.USAGE_MINING, ap_region=1, ap_method=3, an_region=1, an_method=3, identsn=-0.3, diffsn=-0.3, margin_diff=0.0),
          "grel": dict(synth.USAGE_MINING, ap_region=0, ap_method=3, an_region=0, an_method=3, identsn=-0.3, diffsn=-0.3, margin_diff=0.0)}[mname]
ctx = capi.Context(capi.make_config(B, D, sim_precision=prec, grad_chunk_cols=chunk, flags=flags, **mining))
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

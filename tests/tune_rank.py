"""Tuning helper (not a test): per-phase CUDA-event times of ONE RANK of a sharded job, emulated on one GPU through the external-collectives
calls (npair_forward_gathered / npair_row_scalars / npair_backward_gathered): the kernels a rank of `world` runs on its Q x N strip.
   NPAIR_LIB=<variant.so> python tests/tune_rank.py [B] [D] [world] [precision] [flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from npairloss_b200 import capi, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
prec = {"fp16x2": 2, "bf16": 1, "bf16x3": 0}[sys.argv[4] if len(sys.argv) > 4 else "fp16x2"]
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
Q = B // world
x, lab = synth.make_inputs(B, D, 20171230, noise=2.5)
ctx = capi.Context(capi.make_config(Q, D, world=world, rank=0, sim_precision=prec, flags=flags, **synth.USAGE_MINING))
dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
dg = torch.empty((Q, D), dtype=torch.float32, device="cuda")
rs = torch.empty((world, Q, 8), dtype=torch.float32, device="cuda")
def step():
    ctx.forward_gathered(dx, dl)
    ctx.row_scalars(rs[0])
    for r in range(1, world):
        rs[r].copy_(rs[0])                # timing only: the other ranks' records are stand-ins
    ctx.backward_gathered(1.0, rs, dg)
for _ in range(5):
    step()
ctx.profile_enable(True)
acc = np.zeros(9); n = 20
for _ in range(n):
    step(); acc += np.array(ctx.profile_read())
acc /= n
names = ["allg", "prep", "sim", "thr", "row", "build", "grad", "gradT", "bwdx"]
print(os.environ.get("NPAIR_LIB", "default"), sys.argv[1:], f"Q={Q}", " ".join(f"{k}={v*1e3:.1f}us" for k, v in zip(names, acc)), f"sum={acc.sum()*1e3:.1f}us")

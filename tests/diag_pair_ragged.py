"""Diagnostic (GPU): where do the CTA-pair kernels differ from the single-CTA ones on a ragged shape?  python tests/diag_pair_ragged.py [B] [D] [prec]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from npairloss_b200 import capi, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 640
D = int(sys.argv[2]) if len(sys.argv) > 2 else 200
prec = {"bf16x3": capi.PREC_FP32_BF16X3, "bf16": capi.PREC_BF16, "fp16x2": capi.PREC_FP32_FP16X2}[sys.argv[3] if len(sys.argv) > 3 else "bf16x3"]
x, lab = synth.make_inputs(B, D, 20171230, noise=2.5)
dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
out = {}
for name, flags in (("pair", 0), ("1cta", capi.FLAG_SIM_1CTA | capi.FLAG_GRAD_1CTA), ("simpair_grad1", capi.FLAG_GRAD_1CTA), ("sim1_gradpair", capi.FLAG_SIM_1CTA)):
    ctx = capi.Context(capi.make_config(B, D, sim_precision=prec, flags=flags, **synth.USAGE_MINING))
    dg = torch.full_like(dx, float("nan"))
    tops = ctx.forward(dx, dl); ctx.backward(1.0, dg); torch.cuda.synchronize()
    out[name] = dict(tops=np.array(tops, np.float32), S=ctx.debug_read(0, B * B).reshape(B, B),
                     stats=np.stack([ctx.debug_read(w, B) for w in (3, 4, 5, 8, 9)]), grad=dg.cpu().numpy())
    ctx.close()
ref = out["1cta"]
for name in ("pair", "simpair_grad1", "sim1_gradpair"):
    for k in ("tops", "S", "stats", "grad"):
        a, b = out[name][k], ref[k]
        ne = a.view(np.uint32) != b.view(np.uint32)
        msg = f"{name:14s} {k:6s} mismatches={int(ne.sum())}"
        if ne.any():
            idx = np.argwhere(ne)
            msg += f" first={idx[0].tolist()} last={idx[-1].tolist()} a={a[tuple(idx[0])]!r} b={b[tuple(idx[0])]!r}"
            if a.ndim == 2: msg += f" rows={np.unique(idx[:,0])[:6].tolist()}..{np.unique(idx[:,0])[-3:].tolist()} cols={np.unique(idx[:,1])[:6].tolist()}..{np.unique(idx[:,1])[-3:].tolist()}"
        print(msg)

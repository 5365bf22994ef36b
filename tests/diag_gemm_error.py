"""Diagnostic (not a test): where does the tcgen05 similarity error sit?  python tests/diag_gemm_error.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from npairloss_b200 import capi, synth
for (B, D) in [(2048, 512), (2048, 128), (1024, 1024)]:
    x, lab = synth.make_inputs(B, D, 7)
    ref = x.astype(np.float64) @ x.astype(np.float64).T
    xt = torch.from_numpy(x).cuda()
    for prec, nm in [(2, "fp16x2"), (0, "bf16x3"), (1, "bf16")]:
        for be in (0, 1):
            C = capi.debug_gemm(prec, be, xt, xt).cpu().numpy().astype(np.float64)
            err = C - ref
            diag = np.abs(np.diag(err)).max()
            off = np.abs(err - np.diag(np.diag(err)))
            rel = (err / np.maximum(np.abs(ref), 1e-3))
            print(f"B={B} D={D} {nm} backend={be}: diag max {diag:.2e} (mean signed {np.diag(err).mean():+.2e}) offdiag max {off.max():.2e} "
                  f"mean signed rel err on |S|>0.3: {rel[np.abs(ref) > 0.3].mean():+.2e}")

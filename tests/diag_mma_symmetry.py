"""Diagnostic (not a test): is the tcgen05 kind::f16 result invariant under swapping the A and B roles?
 (a) single pass: C1 = A.B^T vs C2 = B.A^T  -> C1 == C2^T bitwise?
 (b) K-interleaved cross terms: one instruction sums 8 products (hi_j*lo_m) and 8 products (lo_j*hi_m); swapping roles permutes the
     16 products inside the instruction.  Equal results <=> the in-instruction sum is order-invariant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from npairloss_b200 import capi
torch.manual_seed(0)
def bf(x): return x.to(torch.bfloat16).to(torch.float32)
for (M, N, K) in [(128, 256, 64), (256, 512, 512), (384, 256, 1024)]:
    A = bf(torch.randn(M, K) * 0.1).cuda(); B = bf(torch.randn(N, K) * 0.1).cuda()
    C1 = capi.debug_gemm(1, 0, A, B); C2 = capi.debug_gemm(1, 0, B, A)
    print(f"(a) {M}x{N}x{K}: bitwise symmetric = {torch.equal(C1, C2.T.contiguous())}, max diff {(C1 - C2.T).abs().max().item():.3e}")
    # (b) build hi/lo pieces of two fp32 matrices, interleave along K in groups of 8
    X = torch.randn(M, K) * 0.1; Y = torch.randn(N, K) * 0.1
    Xh, Yh = bf(X), bf(Y); Xl, Yl = bf(X - Xh), bf(Y - Yh)
    def inter(P, Qm):  # [P(0..7), Q(0..7), P(8..15), Q(8..15), ...]
        r = P.shape[0]
        return torch.stack([P.reshape(r, -1, 8), Qm.reshape(r, -1, 8)], dim=2).reshape(r, -1).contiguous()
    A1, B1 = inter(Xh, Xl).cuda(), inter(Yl, Yh).cuda()      # rank r: x as A operand: [xh, xl] . [yl, yh]
    A2, B2 = inter(Yh, Yl).cuda(), inter(Xl, Xh).cuda()      # rank r': y as A operand: [yh, yl] . [xl, xh]
    D1 = capi.debug_gemm(1, 0, A1, B1); D2 = capi.debug_gemm(1, 0, A2, B2)
    print(f"(b) {M}x{N}x{2*K}: cross-term symmetric = {torch.equal(D1, D2.T.contiguous())}, max diff {(D1 - D2.T).abs().max().item():.3e}, scale {D1.abs().max().item():.3e}")

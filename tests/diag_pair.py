"""Diagnostic (not a test): CTA-pair (cta_group::2) kernels against the single-CTA ones -- S, the fused row statistics and the
   gradient must agree bitwise.  python tests/diag_pair.py [B] [D] [precision] [sim|grad]; runs the configuration twice,
   toggling NPAIR_SIM_1CTA (or NPAIR_GRAD_1CTA)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

def child(B, D, prec, out):
    import torch
    from npairloss_b200 import capi, synth
    x, lab = synth.make_inputs(B, D, 20171230, noise=2.5)
    ctx = capi.Context(capi.make_config(B, D, sim_precision=prec, **synth.USAGE_MINING))
    dx, dl = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
    dg = torch.empty_like(dx)
    tops = ctx.forward(dx, dl); ctx.backward(1.0, dg)
    S = ctx.debug_read(0, B * B)
    st = [ctx.debug_read(w, B) for w in (3, 4, 5, 8, 9)]
    np.savez(out, S=S, st=np.stack(st), g=dg.cpu().numpy(), tops=np.array(tops))
    ctx.profile_enable(True)
    acc = np.zeros(9); n = 10
    for _ in range(n):
        ctx.forward(dx, dl); ctx.backward(1.0, dg); acc += np.array(ctx.profile_read())
    print("sim", "1cta" if os.environ.get("NPAIR_SIM_1CTA") == "1" else "pair", "grad", "1cta" if os.environ.get("NPAIR_GRAD_1CTA") == "1" else "pair",
          B, D, prec, "sim=%.1fus grad=%.1fus" % (acc[2] / n * 1e3, acc[6] / n * 1e3), "loss", tops[0], flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]); sys.exit(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    prec = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    var = "NPAIR_GRAD_1CTA" if (len(sys.argv) > 4 and sys.argv[4] == "grad") else "NPAIR_SIM_1CTA"
    outs = []
    for one in ("1", "0"):
        env = dict(os.environ, **{var: one})
        out = f"/tmp/diag_pair_{one}.npz"
        r = subprocess.run([sys.executable, __file__, "child", str(B), str(D), str(prec), out], env=env, timeout=120)
        if r.returncode != 0: print("child failed", one, r.returncode); sys.exit(1)
        outs.append(np.load(out))
    a, b = outs
    for k in ("S", "st", "g", "tops"):
        same = np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32))
        d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))
        print(k, "bitwise equal" if same else f"DIFFERENT: max abs {d.max():.3e}, n diff {(d > 0).sum()} of {d.size}")

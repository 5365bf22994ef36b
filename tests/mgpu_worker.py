"""Multi-GPU worker (launched by tests/test_multi_gpu.py under torchrun, one rank per GPU, NCCL):
every rank runs npair_forward / npair_backward on its anchor shard; rank 0 checks all ranks against the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from npairloss_b200 import capi, dist_util, synth  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    cfg_flags = int(os.environ.get("NPAIR_TEST_FLAGS", "0"))       # 24 = NCCL exchange instead of peer memory
    for (B, D, mining_name, prec) in [(256 * world, 128, "usage", capi.PREC_FP32_FP16X2), (120 * world, 1024, "usage", capi.PREC_FP32_BF16X3),
                                      (512 * world, 256, "default", capi.PREC_FP32_FP16X2), (64 * world, 96, "local_rel", capi.PREC_FP32_FP16X2)]:
        mining = {"usage": synth.USAGE_MINING, "default": synth.DEFAULT_MINING,
                  "local_rel": dict(synth.DEFAULT_MINING, ap_method=3, an_method=3, identsn=0.0, diffsn=-0.3, margin_diff=-0.01)}[mining_name]
        x, lab = synth.make_inputs(B, D, seed=B + D, noise=2.5)
        Q, rows = dist_util.shard_rows(B, world, rank)
        nid = dist_util.broadcast_bytes(capi.nccl_unique_id() if rank == 0 else None, 128, device=dev)
        ctx = capi.Context(capi.make_config(Q, D, world=world, rank=rank, sim_precision=prec, device=lr, flags=cfg_flags, **mining), nid)
        d_x = torch.from_numpy(np.ascontiguousarray(x[rows])).to(dev)
        d_l = torch.from_numpy(np.ascontiguousarray(lab[rows])).to(dev)
        d_g = torch.full_like(d_x, float("nan"))
        for _ in range(3):                      # several steps: the exchange buffers are double-buffered by step parity
            tops = ctx.forward(d_x, d_l)
            ctx.backward(0.7, d_g)
        tops2 = ctx.forward(d_x, d_l)           # forward-only steps in between (evaluation pattern)
        assert tops2 == tops, (tops2, tops)
        tops = ctx.forward(d_x, d_l)
        ctx.backward(0.7, d_g)
        torch.cuda.synchronize()
        S = torch.from_numpy(ctx.debug_read(0, Q * B).reshape(Q, B)).to(dev)
        ctx.close()
        S_all = [torch.empty_like(S) for _ in range(world)]
        g_all = [torch.empty_like(d_g) for _ in range(world)]
        t_all = [torch.empty(5, device=dev) for _ in range(world)]
        dist.all_gather(S_all, S)
        dist.all_gather(g_all, d_g)
        dist.all_gather(t_all, torch.tensor(tops, device=dev, dtype=torch.float32))
        if rank == 0:
            from oracle import oracle_lib as o
            cfg = o.make_config(Q, D, world=world, faithful_sorts=0, **mining)
            tops_o, dx_o = o.step_world(x, lab, cfg, 0.7, S_inject_all=torch.cat(S_all).cpu().numpy())
            dx = torch.cat(g_all).cpu().numpy()
            tg = torch.stack(t_all).cpu().numpy()
            rel = np.linalg.norm(dx - dx_o) / max(np.linalg.norm(dx_o), 1e-30)
            lerr = np.abs(tg[:, 0] - tops_o[:, 0]).max() / max(np.abs(tops_o[:, 0]).max(), 1e-30)
            good = np.isfinite(dx).all() and rel <= 1e-5 and lerr <= 1e-5 and np.abs(tg[:, 1:4] - tops_o[:, 1:4]).max() * Q <= 1.001
            print(f"[mgpu] flags={cfg_flags} world={world} B={B} D={D} {mining_name} prec={prec}: grad_rel={rel:.2e} loss_rel={lerr:.2e} {'OK' if good else 'FAIL'}", flush=True)
            ok = ok and good
    # ---- world scope (npair_config.global_scope = 1): the sharded job must reproduce the single-rank reference on the
    #      whole batch -- GLOBAL lists over all N x N pairs, loss / gradient normalised by N, identical tops on every rank ----
    for (B, D, mining_name) in [(256 * world, 128, "usage"), (192 * world, 64, "global_rel"), (128 * world, 96, "default")]:
        mining = {"usage": synth.USAGE_MINING, "default": synth.DEFAULT_MINING,
                  "global_rel": dict(synth.USAGE_MINING, ap_region=0, ap_method=4, an_region=0, an_method=3, identsn=-0.4, diffsn=-0.3,
                                     margin_ident=0.01, margin_diff=-0.02)}[mining_name]
        x, lab = synth.make_inputs(B, D, seed=B + D + 1, noise=2.5)
        Q, rows = dist_util.shard_rows(B, world, rank)
        nid = dist_util.broadcast_bytes(capi.nccl_unique_id() if rank == 0 else None, 128, device=dev)
        ctx = capi.Context(capi.make_config(Q, D, world=world, rank=rank, device=lr, flags=cfg_flags, global_scope=1, **mining), nid)
        d_x = torch.from_numpy(np.ascontiguousarray(x[rows])).to(dev)
        d_l = torch.from_numpy(np.ascontiguousarray(lab[rows])).to(dev)
        d_g = torch.full_like(d_x, float("nan"))
        for _ in range(2):
            tops = ctx.forward(d_x, d_l)
            ctx.backward(0.7, d_g)
        torch.cuda.synchronize()
        S = torch.from_numpy(ctx.debug_read(0, Q * B).reshape(Q, B)).to(dev)
        ctx.close()
        S_all = [torch.empty_like(S) for _ in range(world)]
        g_all = [torch.empty_like(d_g) for _ in range(world)]
        t_all = [torch.empty(5, device=dev) for _ in range(world)]
        dist.all_gather(S_all, S); dist.all_gather(g_all, d_g)
        dist.all_gather(t_all, torch.tensor(tops, device=dev, dtype=torch.float32))
        if rank == 0:
            from oracle import oracle_lib as o
            cfg1 = o.make_config(B, D, world=1, faithful_sorts=0, **mining)              # ONE rank holding the whole batch
            tops_o, dx_o = o.step_world(x, lab, cfg1, 0.7, S_inject_all=torch.cat(S_all).cpu().numpy())
            dx = torch.cat(g_all).cpu().numpy()
            tg = torch.stack(t_all).cpu().numpy()
            rel = np.linalg.norm(dx - dx_o) / max(np.linalg.norm(dx_o), 1e-30)
            same_everywhere = bool((tg == tg[0]).all())
            lerr = abs(tg[0, 0] - tops_o[0, 0]) / max(abs(tops_o[0, 0]), 1e-30)
            good = (np.isfinite(dx).all() and rel <= 1e-5 and lerr <= 1e-5 and same_everywhere and np.abs(tg[0, 1:4] - tops_o[0, 1:4]).max() * B <= 1.001
                    and abs(tg[0, 4] - tops_o[0, 4]) <= 2e-6 * abs(tops_o[0, 4]))
            print(f"[mgpu] flags={cfg_flags} GLOBAL SCOPE world={world} B={B} D={D} {mining_name}: grad_rel={rel:.2e} loss_rel={lerr:.2e} "
                  f"tops_identical={same_everywhere} {'OK' if good else 'FAIL'}", flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()

"""world_size-2 gloo tests (CPU): the host-side sharding / exchange logic of the multi-GPU path, with the oracle standing
in for the per-rank device computation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, D, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        from npairloss_b200 import dist_util, synth
        from oracle import oracle_lib as o
        x, lab = synth.make_inputs(B, D, seed=77, noise=2.5)
        Q, rows = dist_util.shard_rows(B, world, rank)
        # unique-id style broadcast
        payload = bytes(range(128)) if rank == 0 else None
        got = dist_util.broadcast_bytes(payload, 128)
        assert got == bytes(range(128))
        # all-gather of the local shard (what ncclAllGather does on the GPU path)
        xl = torch.from_numpy(np.ascontiguousarray(x[rows]))
        ll = torch.from_numpy(np.ascontiguousarray(lab[rows]))
        xs = [torch.empty_like(xl) for _ in range(world)]
        ls = [torch.empty_like(ll) for _ in range(world)]
        dist.all_gather(xs, xl)
        dist.all_gather(ls, ll)
        x_total = torch.cat(xs).numpy()
        lab_total = torch.cat(ls).numpy()
        assert np.array_equal(x_total, x) and np.array_equal(lab_total, lab)
        # per-rank device step stands in: oracle forward + partial backward for this rank
        mining = synth.USAGE_MINING
        cfg = o.make_config(Q, D, world=world, rank=rank, faithful_sorts=0, num_threads=1, **mining)
        L = o.lib()
        buf = np.zeros(L.npo_state_floats(C.byref(cfg)), dtype=np.float32)
        st = o.NpoState()
        L.npo_state_bind(C.byref(cfg), o._fp(buf), C.byref(st))
        tops = np.zeros(5, np.float32)
        assert L.npo_forward(C.byref(cfg), o._fp(x_total), o._fp(lab_total), None, C.byref(st), o._fp(tops)) == 0
        local = np.zeros((Q, D), np.float32)
        total = np.zeros((B, D), np.float32)
        assert L.npo_backward_partial(C.byref(cfg), o._fp(x_total), C.byref(st), C.c_float(1.0), o._fp(local), o._fp(total)) == 0
        # fold the reference's 1/2 and 1/k into the halves exactly like npair_backward_partial does
        local_half = 0.5 * local
        total_half = torch.from_numpy((0.5 / world) * total)
        dist.all_reduce(total_half)                      # reduce-scatter == all-reduce + own slice (.cu:467-497)
        dx = dist_util.blend_reference(local_half, total_half.numpy(), rows)
        np.save(os.path.join(outdir, f"dx{rank}.npy"), dx)
        np.save(os.path.join(outdir, f"tops{rank}.npy"), tops)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_gloo_sharded_step_matches_single_process(tmp_path, oracle, world):
    from npairloss_b200 import synth
    B, D = 64, 24
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, D, str(tmp_path)), nprocs=world, join=True)
    x, lab = synth.make_inputs(B, D, seed=77, noise=2.5)
    tops_ref, dx_ref = oracle.step_world(x, lab, oracle.make_config(B // world, D, world=world, faithful_sorts=0, **synth.USAGE_MINING), 1.0)
    dx = np.concatenate([np.load(tmp_path / f"dx{r}.npy") for r in range(world)])
    tops = np.stack([np.load(tmp_path / f"tops{r}.npy") for r in range(world)])
    np.testing.assert_allclose(tops, tops_ref, rtol=1e-6)
    assert np.linalg.norm(dx - dx_ref) <= 2e-6 * np.linalg.norm(dx_ref)


def test_shard_rows():
    from npairloss_b200 import dist_util
    assert dist_util.shard_rows(8192, 8, 3) == (1024, slice(3072, 4096))
    with pytest.raises(ValueError):
        dist_util.shard_rows(10, 4, 0)

"""Host-side helpers for the anchor-sharded multi-GPU path (one process per GPU, torch.distributed for plumbing).
Pure host logic: exercised on CPU with the gloo backend (tests/test_dist_cpu.py) and on GPUs with NCCL (bench.py)."""
from __future__ import annotations

import numpy as np


def shard_rows(B: int, world: int, rank: int):
    """Contiguous anchor block of rank `rank` (labels travel with their rows; self pair = column i + rank*Q,
    reference npair_multi_class_loss.cu:54)."""
    if B % world:
        raise ValueError(f"global batch {B} is not divisible by world size {world}")
    Q = B // world
    return Q, slice(rank * Q, (rank + 1) * Q)


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """Ships `nbytes` bytes (e.g. the 128-byte NCCL unique id, which the reference's MPI fork would MPI_Bcast) from
    rank `src` to every rank through the default process group."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        assert payload is not None and len(payload) == nbytes
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def blend_reference(local_half: np.ndarray, total_halves_sum: np.ndarray, rows: slice) -> np.ndarray:
    """bottom.diff of one rank from the two pieces of npair_backward_partial: local_half + (sum over ranks of
    total_half)[own rows]  (reference .cu:462-497 with the 1/2 and 1/k factors already folded into the halves)."""
    return local_half + total_halves_sum[rows]

"""Frame sharding for one-process-per-GPU runs.

Frames are independent when the seeds are None, the forest is replicated on every GPU, and there
is no exchange step: each rank predicts a contiguous block of frames and the 56-byte results are
gathered on the host.  No collective touches the data path (NCCL is not used for it).
"""
from __future__ import annotations

import numpy as np

from . import capi


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`: ceil(n/world) frames per rank, the tail ranks may
    get fewer (or none).  Contiguous blocks keep a sequence's frames together."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = (n_frames + world - 1) // world
    lo = min(rank * per, n_frames)
    return lo, min(lo + per, n_frames)


def gather_results(local: np.ndarray, n_frames: int, dist=None) -> np.ndarray | None:
    """Host gather of the per-rank result blocks into one [n_frames] array on rank 0 (None on the
    other ranks).  `dist` is torch.distributed (any backend) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert len(local) == n_frames
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local.tobytes(), parts, dst=0)
    if rank != 0:
        return None
    out = np.zeros(n_frames, capi.RESULT_DTYPE)
    for r, blob in enumerate(parts):
        lo, hi = shard_range(n_frames, r, world)
        out[lo:hi] = np.frombuffer(blob, capi.RESULT_DTYPE)
    return out

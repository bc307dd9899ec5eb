"""Host-side multi-GPU logic on CPU: contiguous frame sharding + host gather, exercised with a
world_size-2 gloo group.  The per-shard compute stand-in is the oracle (tests may call it); the
product's kernels are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest

from depthhead_b200 import capi, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 15000, 1024):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, hi = shard.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                got.extend(range(lo, hi))
            assert got == list(range(n))
    assert shard.shard_range(15000, 7, 8) == (13125, 15000)
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_frames_depend_only_on_seed_and_index():
    a = synth.make_frames(6, seed=5)
    b = synth.make_frames(3, seed=5, start_index=3)
    assert np.array_equal(a[3:], b)
    s = synth.make_frames(4, seed=5, sequence=True, start_index=623)
    t = synth.make_frames(2, seed=5, sequence=True, start_index=625)
    assert np.array_equal(s[2:], t)


def _worker(rank, world, port, n_frames, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    arr = synth.make_forest(seed=3, n_trees=3, max_depth=5)
    of = oracle.OracleForest(arr, 20, 80, 80, 8.0, 10)
    lo, hi = shard.shard_range(n_frames, rank, world)
    frames = synth.make_frames(hi - lo, seed=9, start_index=lo)
    mid, rot, _ = of.predict_batch(frames, synth.KINECT_K, mode=oracle.MODE_SAT) if hi > lo else (np.zeros((0, 3), np.float32), np.zeros((0, 3)), 0)
    local = np.zeros(hi - lo, capi.RESULT_DTYPE)
    local["mid_point"], local["rotation"] = mid, rot
    full = shard.gather_results(local, n_frames, dist)
    if rank == 0:
        q.put((full["mid_point"].copy(), full["rotation"].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_process():
    import torch.multiprocessing as mp
    import oracle
    n_frames = 5
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    mid, rot = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    arr = synth.make_forest(seed=3, n_trees=3, max_depth=5)
    of = oracle.OracleForest(arr, 20, 80, 80, 8.0, 10)
    m1, r1, _ = of.predict_batch(synth.make_frames(n_frames, seed=9), synth.KINECT_K, mode=oracle.MODE_SAT)
    assert np.array_equal(mid, m1) and np.array_equal(rot, r1)  # byte-identical to one process

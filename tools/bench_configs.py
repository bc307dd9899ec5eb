"""Stage times of BASELINE.json configs[3] (50 sparse trees of depth 20, stride 1) and configs[4]
(vote-heavy: 32..128 votes per leaf, stride 1) on one GPU — parity-test configurations, measured
here only to see which kernel binds them (the bench line is configs[1]).

    python tools/bench_configs.py [--frames 32]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from depthhead_b200 import Context, HoughPrediction, IntrinsicMatrix, synth  # noqa: E402


def run(name, arr, stride, frames, ctx, K, steps=3):
    hp = HoughPrediction.from_arrays(arr, stepwidth=stride)
    dev = torch.from_numpy(frames.view(np.int16)).cuda()
    n, h, w = frames.shape

    def step():
        return hp.predict_batch(None, K, ctx=ctx, device_ptr=dev.data_ptr(), n=n, w=w, h=h)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    ctx.enable_stage_timing(True)
    step()
    st = ctx.stage_ms()
    cnt = ctx.counters()
    ctx.enable_stage_timing(False)
    visits, evals = cnt["node_visits"], cnt["evals"]
    trav_s = st["traverse"] / 1000.0
    out = {"config": name, "frames": n, "frames_per_s": n / dt, "ms_per_frame": 1000.0 * dt / n,
           "stage_ms_per_frame": {k: v / n for k, v in st.items()},
           "nodes": hp.n_nodes, "node_table_MB": hp.n_nodes * 16 / 1e6, "evals_per_frame": evals / n,
           "mean_visited_depth": visits / max(1, evals), "votes_per_frame": (cnt["centre_votes"] + cnt["rot_votes"]) / n,
           "traverse_node_visits_per_s": visits / trav_s if trav_s > 0 else None,
           "traverse_algorithmic_GBps": (visits * 56 + evals * 16) / trav_s / 1e9 if trav_s > 0 else None}
    print(json.dumps(out))
    hp.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--trained", action="store_true", help="also a forest trained on synthetic frames (dh_train_learn)")
    ap.add_argument("--deep", action="store_true", help="also a 50-tree depth-20 forest with ~2^17 nodes per tree (node tables ~ L2 size)")
    a = ap.parse_args()
    K = IntrinsicMatrix.default_kinect_intrinsic()
    ctx = Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    frames = synth.make_frames(a.frames, seed=8)
    run("configs[3]: 50 sparse trees depth 20, stride 1", synth.make_forest(seed=9, n_trees=50, max_depth=20, stop_prob=0.3), 1, frames, ctx, K)
    if a.deep:
        t0 = time.perf_counter()
        arr = synth.make_forest(seed=9, n_trees=50, max_depth=20, stop_prob=0.11)
        sys.stderr.write("deep forest generated in %.1f s\n" % (time.perf_counter() - t0))
        run("configs[3] deep: 50 trees depth 20 with stop probability 0.11, stride 1", arr, 1, frames[:8], ctx, K, steps=2)
        del arr
    run("configs[4]: vote-heavy, 10 trees depth 8, 32..128 votes per leaf, stride 1",
        synth.make_forest(seed=13, n_trees=10, max_depth=8, votes_lo=32, votes_hi=128), 1, frames, ctx, K)
    run("configs[1] forest at stride 10 (the reference's real-time setting)", synth.make_forest(seed=1, n_trees=10, max_depth=15), 10,
        synth.make_frames(512, seed=2024), ctx, K)
    if a.trained:
        # a forest TRAINED on synthetic frames with ground truth (dh_train_learn, the reference trainer's defaults but 500
        # candidates per node), then the configs[1] workload and its accuracy on held-out frames
        from depthhead_b200 import train
        tf, tc, tr, tm = synth.make_frames(300, seed=101, with_truth=True)
        data = [dict(depth=tf[i], mask=tm[i], intrinsic=K, pos3d=tc[i], rot=tr[i]) for i in range(len(tf))]
        hl = train.HoughLearning(10, 80, 80, 15, 10, 5200, 0.3, 500, 20, 5.0)
        t0 = time.perf_counter()
        hp = hl.learn_native(8.0, data, seed=1, ctx=ctx)
        t_train = time.perf_counter() - t0
        hp.stepwidth = 5
        frames2, centres2, _, _ = synth.make_frames(512, seed=2024, with_truth=True)
        dev = torch.from_numpy(frames2.view(np.int16)).cuda()
        for _ in range(2):
            out = hp.predict_batch(None, K, ctx=ctx, device_ptr=dev.data_ptr(), n=512, w=640, h=480)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = hp.predict_batch(None, K, ctx=ctx, device_ptr=dev.data_ptr(), n=512, w=640, h=480)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        ctx.enable_stage_timing(True)
        hp.predict_batch(None, K, ctx=ctx, device_ptr=dev.data_ptr(), n=512, w=640, h=480)
        st, cnt = ctx.stage_ms(), ctx.counters()
        ctx.enable_stage_timing(False)
        d = out["mid_point"].astype(np.float64) - centres2
        print(json.dumps({"config": "forest trained on 300 synthetic frames (10 trees, depth <= 15, 5200 samples per tree, 500 candidates per node), "
                                    "stride 5, 512 held-out frames",
                          "train_seconds": t_train, "nodes": hp.n_nodes, "leaves": hp.n_leaves, "votes": hp.n_votes,
                          "frames_per_s": 512 / dt, "stage_ms_per_frame": {k: v / 512 for k, v in st.items()},
                          "mean_visited_depth": cnt["node_visits"] / max(1, cnt["evals"]),
                          "votes_per_frame": (cnt["centre_votes"] + cnt["rot_votes"]) / 512, "cube_rebuilds": cnt["cube_rebuilds"],
                          "meanshift_rounds_per_accumulator": cnt["meanshift_iters"] / 1024.0,
                          "median_lateral_error_mm": float(np.median(np.hypot(d[:, 0], d[:, 1]))),
                          "median_depth_error_mm": float(np.median(d[:, 2]))}))


if __name__ == "__main__":
    main()

"""Training throughput at the reference's scale (hough_tree_trainer.rs defaults: 5200 samples per
tree, 2000 candidate features per node, depth 15): seconds per tree on the GPU scorer, and the CPU
restatement of the same callbacks (naive binarize + impurity per candidate) on a sample.

    python tools/bench_train.py [--frames 300] [--trees 1]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from depthhead_b200 import Context, IntrinsicMatrix, synth, train  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--trees", type=int, default=1)
    ap.add_argument("--features", type=int, default=2000)
    ap.add_argument("--subset", type=int, default=5200)
    ap.add_argument("--depth", type=int, default=15)
    a = ap.parse_args()
    K = IntrinsicMatrix.default_kinect_intrinsic()
    ctx = Context(0)
    frames, centres, rots, masks = synth.make_frames(a.frames, seed=101, with_truth=True)
    data = [dict(depth=frames[i], mask=masks[i], intrinsic=K, pos3d=centres[i], rot=rots[i]) for i in range(a.frames)]
    hl = train.HoughLearning(10, 80, 80, a.depth, a.trees, a.subset, 0.3, a.features, 20, 5.0)
    # time the scorer calls separately from the host loop
    t_score = [0.0]
    cand_samples = [0]
    orig = train.TrainSet.score_level

    def timed(self, idx, off, rects, thr, depth, steep):
        t0 = time.perf_counter()
        out = orig(self, idx, off, rects, thr, depth, steep)
        t_score[0] += time.perf_counter() - t0
        cand_samples[0] += int(np.sum(np.diff(np.asarray(off).astype(np.int64)) * (np.asarray(thr).size // (len(off) - 1))))
        return out
    train.TrainSet.score_level = timed
    t0 = time.perf_counter()
    hp = hl.learn(8.0, data, seed=7, ctx=ctx)
    t_learn = time.perf_counter() - t0
    # the same trees grown by the C++ loop behind dh_train_forest (samples already extracted)
    t0 = time.perf_counter()
    hp_native = hl.train_native(8.0, *hl.last_samples, seed=8, ctx=ctx)
    t_native = time.perf_counter() - t0
    assert hp_native.n_nodes == hp.n_nodes
    # CPU restatement on a sample: root-like node, a few candidates
    rng = np.random.default_rng(3)
    P, F, O, R = [], [], [], []
    for i in range(min(a.frames, 130)):
        org, flag, offs, rr = train.extract_samples(frames[i], masks[i], K, centres[i], rots[i], 10, 80, 80)
        for part in (rng.permutation(np.flatnonzero(flag == 0))[:20], rng.permutation(np.flatnonzero(flag != 0))[:20]):
            for j in part:
                x0, y0 = org[j]
                P.append(frames[i][y0:y0 + 80, x0:x0 + 80]); F.append(flag[j]); O.append(offs[j]); R.append(rr[j])
    P, F, O, R = np.stack(P), np.asarray(F, np.uint8), np.asarray(O, np.float32), np.asarray(R, np.float64)
    idx = np.arange(len(P), dtype=np.uint32)
    rects, thr = hl.param_set(rng, 8)
    thr[:] = rng.uniform(-30, 30, 8)
    t0 = time.perf_counter()
    for j in range(8):
        bits = oracle.train_binarize(P, idx, rects[j], thr[j])
        le, ri = idx[bits == 0], idx[bits != 0]
        if len(le) and len(ri):
            oracle.train_impurity(F, O, R, le, ri, 3, 5.0)
    t_cpu = time.perf_counter() - t0
    cpu_rate = 8 * len(P) / t_cpu
    print(json.dumps({
        "trees": a.trees, "samples_in_set": int(40 * a.frames), "subset_per_tree": a.subset, "features_per_node": a.features,
        "max_depth": a.depth, "nodes": int(hp.n_nodes), "leaves": int(hp.n_leaves),
        "learn_seconds": t_learn, "seconds_per_tree": t_learn / a.trees, "scorer_seconds": t_score[0],
        "native_trainer_seconds": t_native, "native_seconds_per_tree": t_native / a.trees,
        "candidate_x_sample_evaluations": cand_samples[0], "gpu_candidate_samples_per_s": cand_samples[0] / t_score[0],
        "cpu_port_candidate_samples_per_s_1_thread": cpu_rate,
        "cpu_port_seconds_per_tree_1_thread_extrapolated": cand_samples[0] / a.trees / cpu_rate,
        "note": "learn_seconds: Python mirror incl. sample extraction from the frames; native_trainer_seconds: dh_train_forest (C++ loop) on the extracted samples incl. box tables and flattening"}))


if __name__ == "__main__":
    main()

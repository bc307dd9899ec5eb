"""Regenerates tests/golden/oracle_small.npz.

The reference is pure Rust with an un-vendored dependency (stamm 0.2.0) and there is no Rust
toolchain in the build image, so these vectors are produced by the C++ ORACLE (not by the reference
binary): they pin the oracle against silent regressions and give the CUDA path a committed fixture.
Inputs are regenerated from seeds, so only the outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from depthhead_b200 import synth  # noqa: E402

CASES = [  # name, forest kwargs, stepwidth, frame seed, n frames
    ("d6t4s10", dict(seed=3, n_trees=4, max_depth=6), 10, 11, 3),
    ("d9t6s7", dict(seed=13, n_trees=6, max_depth=9, shuffle_nodes=True), 7, 12, 2),
    ("sparse", dict(seed=5, n_trees=5, max_depth=11, stop_prob=0.25, ragged_rects=True), 6, 13, 2),
]


def main():
    out = {}
    for name, fk, step, fseed, n in CASES:
        arr = synth.make_forest(**fk)
        of = oracle.OracleForest(arr, step, 80, 80, 8.0, 20)
        frames = synth.make_frames(n, seed=fseed)
        for i, d in enumerate(frames):
            tr = of.predict(d, synth.KINECT_K, mode=oracle.MODE_NAIVE, keep=True)
            p = "%s/%d/" % (name, i)
            out[p + "leaf"] = tr.leaf
            out[p + "gate"] = tr.gate
            out[p + "guess_pos"] = tr.guess_pos
            out[p + "guess_rot"] = tr.guess_rot
            out[p + "seed_mid"] = tr.seed_mid
            out[p + "seed_rot"] = tr.seed_rot
            out[p + "mid_point"] = tr.mid_point
            out[p + "rotation"] = tr.rotation
            out[p + "n_votes"] = np.array([tr.n_mid_votes, tr.n_rot_votes], np.int64)
            out[p + "ms_mid"] = tr.ms_mid
            out[p + "ms_rot"] = tr.ms_rot
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()

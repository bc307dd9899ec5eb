"""Two independent restatements of the reference's prediction path must agree bit for bit:
oracle/depthhead_oracle.cpp (C++, tables, the checker of every GPU test) and
oracle/second_opinion.py (written from the Rust source again, pure Python with numpy scalars of the
reference's widths, dict-backed accumulators).  This removes the single-reader risk from the
oracle; it does not pin either to the Rust binary (stamm's part stays unpinned in both)."""
import numpy as np
import pytest

import oracle
from oracle import second_opinion as so
from depthhead_b200 import synth


def _crop(frame, h, w):
    y0, x0 = (480 - h) // 2 - 40, (640 - w) // 2
    return np.ascontiguousarray(frame[y0:y0 + h, x0:x0 + w])


CASES = [
    dict(forest=dict(seed=3, n_trees=3, max_depth=5), stride=9, hw=(150, 190), sigma=8.0, iters=6, guess=None),
    dict(forest=dict(seed=6, n_trees=4, max_depth=6, stop_prob=0.25, ragged_rects=True), stride=11, hw=(140, 200), sigma=5.0, iters=4, guess=None),
    dict(forest=dict(seed=8, n_trees=2, max_depth=4, tie_thresholds=True), stride=13, hw=(130, 170), sigma=8.0, iters=3,
         guess=([12.5, -40.25, 950.0], [0.2, -0.35, 0.05])),
]


@pytest.mark.parametrize("case", CASES, ids=["full", "sparse-ragged", "ties-seeded"])
def test_cpp_oracle_equals_the_second_restatement(case):
    arr = synth.make_forest(**case["forest"])
    frame = _crop(synth.make_frames(1, seed=41)[0], *case["hw"])
    mg, rg = case["guess"] if case["guess"] else (None, None)
    of = oracle.OracleForest(arr, case["stride"], 80, 80, case["sigma"], case["iters"])
    tr = of.predict(frame, synth.KINECT_K, mg, rg, mode=oracle.MODE_NAIVE, keep=True)
    b = so.predict(arr, frame, synth.KINECT_K, case["stride"], 80, 80, case["sigma"], case["iters"], mg, rg)
    assert b["leaf"].shape == tr.leaf.shape and np.array_equal(b["leaf"], tr.leaf)
    assert np.array_equal(b["gate"], tr.gate.astype(bool))
    g = b["gate"]
    assert np.array_equal(b["p3"][g].view(np.uint32), tr.p3[g].view(np.uint32))
    assert np.array_equal(b["guess_pos"], tr.guess_pos)
    dense = np.zeros(8000, np.uint32)
    for (x, y, z), v in b["guess_rot"].items():
        dense[z * 400 + y * 20 + x] = v
    assert np.array_equal(dense, tr.guess_rot)
    assert tuple(tr.seed_mid) == tuple(b["seed_mid"]) and tuple(tr.seed_rot) == tuple(b["seed_rot"])
    for mine, keys, vals in ((b["mid"], tr.mid_keys, tr.mid_vals), (b["rot"], tr.rot_keys, tr.rot_vals)):
        theirs = {tuple(int(c) for c in k): int(v) for k, v in zip(keys, vals)}
        # a zero-weight vote inserts a key with value 0 in the reference's map; mean-shift skips it (meanshift.rs:364-366)
        assert {k: v for k, v in mine.items() if v} == {k: v for k, v in theirs.items() if v}
    assert [tuple(p) for p in tr.ms_mid] == b["ms_mid"] and [tuple(p) for p in tr.ms_rot] == b["ms_rot"]
    assert np.array_equal(b["mid_point"], tr.mid_point) and np.array_equal(b["rotation"].view(np.uint64), tr.rotation.view(np.uint64))
    assert g.sum() > 0 and len(b["mid"]) > 0, "the case must exercise the vote path"

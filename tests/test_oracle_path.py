"""CPU tests of the oracle's prediction path: internal consistency, reference-derived properties
 and the committed golden vectors."""
import importlib.util
import os

import numpy as np
import pytest

import oracle
from depthhead_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_cases():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.CASES


def test_golden_vectors():
    g = np.load(os.path.join(HERE, "golden", "oracle_small.npz"))
    for name, fk, step, fseed, n in _load_cases():
        arr = synth.make_forest(**fk)
        of = oracle.OracleForest(arr, step, 80, 80, 8.0, 20)
        for i, d in enumerate(synth.make_frames(n, seed=fseed)):
            tr = of.predict(d, synth.KINECT_K, mode=oracle.MODE_SAT, keep=True)
            p = "%s/%d/" % (name, i)
            for key in ("leaf", "gate", "guess_pos", "guess_rot", "seed_mid", "seed_rot", "mid_point", "rotation",
                        "ms_mid", "ms_rot"):
                assert np.array_equal(getattr(tr, key), g[p + key]), (name, i, key)
            assert [tr.n_mid_votes, tr.n_rot_votes] == g[p + "n_votes"].tolist()


def test_naive_equals_sat_and_threads(small_case):
    arr, js, frames = small_case
    of = oracle.OracleForest.from_json(js)
    for d in frames:
        a = of.predict(d, synth.KINECT_K, mode=oracle.MODE_NAIVE, tree_threads=1)
        b = of.predict(d, synth.KINECT_K, mode=oracle.MODE_SAT, tree_threads=1)
        c = of.predict(d, synth.KINECT_K, mode=oracle.MODE_NAIVE, tree_threads=3)
        for x in (b, c):
            assert np.array_equal(a.leaf, x.leaf)
            assert np.array_equal(a.mid_keys, x.mid_keys) and np.array_equal(a.mid_vals, x.mid_vals)
            assert np.array_equal(a.rot_keys, x.rot_keys) and np.array_equal(a.rot_vals, x.rot_vals)
            assert np.array_equal(a.mid_point, x.mid_point) and np.array_equal(a.rotation, x.rotation)


def test_rect_average_naive_vs_sat_and_empty_rect():
    import ctypes as C
    rng = np.random.default_rng(0)
    img = rng.integers(0, 65536, (120, 160)).astype(np.uint16)
    L = oracle.lib()
    sub = np.array([17, 9, 80, 80], np.uint32)
    for _ in range(200):
        x0, y0 = rng.integers(0, 80, 2)
        x1, y1 = rng.integers(x0, 81), rng.integers(y0, 81)
        r = np.array([x0, y0, x1, y1], np.uint32)
        a = L.orc_rect_average(img.ctypes.data_as(C.c_void_p), 160, 120, sub.ctypes.data_as(C.c_void_p),
                               r.ctypes.data_as(C.c_void_p), 0)
        b = L.orc_rect_average(img.ctypes.data_as(C.c_void_p), 160, 120, sub.ctypes.data_as(C.c_void_p),
                               r.ctypes.data_as(C.c_void_p), 1)
        assert a == b
        if x1 == x0 or y1 == y0:
            assert a == 0.0  # count == 0 -> 0.0 (types.rs:335-337)
        else:
            ref = img[9 + y0:9 + y1, 17 + x0:17 + x1].astype(np.uint64).sum() / float((x1 - x0) * (y1 - y0))
            assert a == ref


def test_sliding_window_counts():
    arr = synth.make_forest(seed=1, n_trees=1, max_depth=2)
    for step, (npx, npy) in ((10, (56, 40)), (5, (112, 80)), (1, (560, 400)), (7, (80, 58))):
        if step == 1:
            continue  # 224 000 patches: covered on the GPU
        of = oracle.OracleForest(arr, step, 80, 80, 8.0, 1)
        tr = of.predict(np.zeros((480, 640), np.uint16), synth.KINECT_K)
        assert (tr.npx, tr.npy) == (npx, npy)  # P = 2240 (s=10), 8960 (s=5)


def test_background_and_degenerate_frames(small_case):
    arr, js, frames = small_case
    of = oracle.OracleForest.from_json(js)
    tr = of.predict(np.zeros((480, 640), np.uint16), synth.KINECT_K)
    assert tr.valid.sum() == 0 and np.all(tr.leaf == -1)
    # all-zero grids -> centre seed from cell (0,0) with mean depth 0 -> (0,0,0); rotation seed
    # cell (0,0,0) -> 9 degrees -> bin 3
    assert tr.seed_mid.tolist() == [0, 0, 0] and tr.seed_rot.tolist() == [3, 3, 3]
    assert tr.ms_mid_zero and tr.ms_rot_zero
    assert tr.mid_point.tolist() == [0.0, 0.0, 0.0]
    assert np.allclose(tr.rotation, (3 - 60) / 60 * 3.14159)
    # a single valid pixel makes every patch containing it non-background (prediction.rs:567-571)
    one = np.zeros((480, 640), np.uint16)
    one[240, 320] = 1000
    tr = of.predict(one, synth.KINECT_K)
    assert tr.valid.sum() == 8 * 8  # stride 10, 80x80 patches


def test_single_vote_and_zero_prob_leaves_never_vote():
    # one tree that is a single leaf (no nodes): every non-background patch lands in it
    def forest(prob, offs, rots):
        return dict(n_trees=1, tree_node_off=np.array([0, 0]), tree_leaf_off=np.array([0, 1]),
                    rects=np.zeros((0, 8), np.int64), threshold=np.zeros(0), child=np.zeros((0, 2), np.int32),
                    prob=np.array([prob]), vote_off=np.array([0, len(offs)]),
                    offsets=np.asarray(offs, np.float32).reshape(-1, 3), rotations=np.asarray(rots, np.float64).reshape(-1, 3))
    d = synth.make_frames(1, seed=3)[0]
    one = oracle.OracleForest(forest(0.9, [[1, 2, 3]], [[4, 5, 6]]), 10, 80, 80, 8.0, 5)
    tr = one.predict(d, synth.KINECT_K)
    assert tr.gate.sum() > 0 and tr.n_mid_votes == 0 and tr.n_rot_votes == 0  # NaN trace (n == 1)
    two = oracle.OracleForest(forest(0.9, [[1, 2, 3], [2, 3, 4]], [[4, 5, 6], [5, 6, 7]]), 10, 80, 80, 8.0, 5)
    tr = two.predict(d, synth.KINECT_K)
    assert tr.n_rot_votes == 2 * int(tr.gate.sum()) and tr.n_mid_votes > 0
    assert set(np.unique(tr.mid_vals)) <= {450 * k for k in range(1, 200)}  # valtoadd = 900 / 2
    low = oracle.OracleForest(forest(0.7, [[1, 2, 3], [2, 3, 4]], [[4, 5, 6], [5, 6, 7]]), 10, 80, 80, 8.0, 5)
    assert low.predict(d, synth.KINECT_K).gate.sum() == 0  # strict > 0.7 (prediction.rs:584)


def test_rotation_bins_wrap_and_truncate():
    def forest(rots):
        n = len(rots)
        return dict(n_trees=1, tree_node_off=np.array([0, 0]), tree_leaf_off=np.array([0, 1]),
                    rects=np.zeros((0, 8), np.int64), threshold=np.zeros(0), child=np.zeros((0, 2), np.int32),
                    prob=np.array([1.0]), vote_off=np.array([0, n]), offsets=np.zeros((n, 3), np.float32),
                    rotations=np.asarray(rots, np.float64))
    d = np.zeros((480, 640), np.uint16)
    d[200:280, 280:360] = 900
    # -1.5 deg -> (-0.5) as i32 = 0 -> bin 60 (truncation, not floor); 179.9 -> 59+60 = 119;
    # 181 -> 60+60 = 120 -> wraps to 0; -181 -> -60+60 = 0
    # (two identical votes per leaf: covariance 0 passes the trace gate)
    of = oracle.OracleForest(forest([[-1.5, 179.9, 181.0]] * 2), 40, 80, 80, 8.0, 1)
    tr = of.predict(d, synth.KINECT_K)
    assert tr.rot_keys.tolist() == [[60, 119, 0]]
    assert tr.guess_rot[0 * 400 + 19 * 20 + 10] == tr.rot_vals[0] > 0
    of = oracle.OracleForest(forest([[2.9, -179.9, -181.0]] * 2), 40, 80, 80, 8.0, 1)
    tr = of.predict(d, synth.KINECT_K)
    assert tr.rot_keys.tolist() == [[60, 1, 0]]  # 2.9 -> 0 -> 60; -179.9 -> -59 -> 1; -181 -> -60 -> 0


def test_caller_seeds_and_result_quantum(small_case):
    arr, js, frames = small_case
    of = oracle.OracleForest.from_json(js)
    tr = of.predict(frames[0], synth.KINECT_K, midp_guess=[10.9, -20.9, 900.5], rot_guess=[0.1, -0.1, 0.0])
    assert tr.seed_mid.tolist() == [10, -20, 900]  # `as i32` truncation (prediction.rs:438)
    exp = [int((g * 180.0 / 3.14159 + 180.0) * 120.0 / 360.0) for g in (0.1, -0.1, 0.0)]
    assert tr.seed_rot.tolist() == exp
    # outputs are integer-valued mm and multiples of 3 degrees in radians(3.14159)
    assert np.all(tr.mid_point == np.round(tr.mid_point))
    bins = tr.rotation / 3.14159 * 60 + 60
    assert np.allclose(bins, np.round(bins), atol=1e-9)


def test_mask_and_hough_image_modes_agree(small_case):
    arr, js, frames = small_case
    of = oracle.OracleForest.from_json(js)
    d = frames[0]
    assert np.array_equal(of.predict_mask(d, mode=0), of.predict_mask(d, mode=1))
    assert np.array_equal(of.hough_image_raw(d, synth.KINECT_K, mode=0), of.hough_image_raw(d, synth.KINECT_K, mode=1))
    assert of.predict_mask(d).max() > 0


def test_image_smaller_than_patch_is_an_error(small_case):
    arr, js, frames = small_case
    of = oracle.OracleForest.from_json(js)
    with pytest.raises(RuntimeError):
        of.predict(np.zeros((60, 60), np.uint16), synth.KINECT_K)

"""Training callbacks on the CPU: the oracle's restatement of
HoughTreeFunctions::{impurity, early_stop} (src/hough/houghforest.rs:250-311) against values worked
out by hand, and the host-side pieces of depthhead_b200.train (param_set / scale_and_replace,
sample extraction) against the oracle and the reference's own known answers."""
import math

import numpy as np
import pytest

import oracle
from depthhead_b200 import IntrinsicMatrix, synth, train


def _ln(x):
    return 0.0 if x == 0.0 else math.log(x)


def test_impurity_by_hand():
    # 6 samples: objects 0,1,2,4; left = [0,1,3], right = [2,4,5]
    is_obj = np.array([1, 1, 1, 0, 1, 0], np.uint8)
    off = np.array([[1, 2, 3], [3, 1, 0], [0, 0, 1], [9, 9, 9], [2, 5, 1], [7, 7, 7]], np.float32)
    rot = np.array([[10, 0, 5], [20, 5, 5], [0, 1, 2], [8, 8, 8], [4, 3, 1], [6, 6, 6]], np.float64)
    left, right = [0, 1, 3], [2, 4, 5]
    v, st, bad = oracle.train_impurity(is_obj, off, rot, left, right, depth=3, steepness=5.0)
    assert not bad and list(st["n"]) == [3, 3] and list(st["n_pos"]) == [2, 2]

    def entropy(p):
        return p * _ln(p) + (1 - p) * _ln(1 - p)

    def cov_det(v):  # two points: covariance of rank one, determinant exactly 0 up to rounding
        m = (v[0] + v[1]) / 2.0
        c = sum(np.outer(x - m, x - m) for x in v) / 1.0
        return (c[0, 0] * (c[1, 1] * c[2, 2] - c[1, 2] * c[2, 1]) - c[1, 0] * (c[0, 1] * c[2, 2] - c[0, 2] * c[2, 1])
                + c[2, 0] * (c[0, 1] * c[1, 2] - c[0, 2] * c[1, 1]))
    dl = cov_det(off[[0, 1]].astype(np.float64)) + cov_det(rot[[0, 1]])
    dr = cov_det(off[[2, 4]].astype(np.float64)) + cov_det(rot[[2, 4]])
    assert st["det_off"][0] + st["det_rot"][0] == dl and st["det_off"][1] + st["det_rot"][1] == dr
    reg = 0.5 * (math.log(dl) if dl > 0 else 0.0) + 0.5 * (math.log(dr) if dr > 0 else 0.0)
    want = -(0.5 * entropy(2 / 3) + 0.5 * entropy(2 / 3)) + (1 - math.exp(-(3 / 5.0))) * reg
    assert v == want
    assert oracle.train_impurity_from_stats(st, 3, 5.0) == v
    # a side without objects contributes 0 to the regression term, and ln!(0) = 0 in its entropy
    v2, st2, _ = oracle.train_impurity(is_obj, off, rot, [3, 5], [0, 1, 2, 4], depth=0, steepness=5.0)
    assert list(st2["n_pos"]) == [0, 4] and math.isnan(st2["det_off"][0])
    assert v2 == -(2 / 6 * 0.0 + 4 / 6 * 0.0) + (1 - math.exp(-0.0)) * 0.0   # depth 0: the weight 1 - e^0 is 0
    # a single object on a side: 0/0 covariance -> NaN determinant -> neither match arm -> 0
    v3, st3, _ = oracle.train_impurity(is_obj, off, rot, [0, 3], [1, 2, 4, 5], depth=2, steepness=5.0)
    assert st3["n_pos"][0] == 1 and math.isnan(st3["det_off"][0])
    assert math.isfinite(v3)


def test_early_stop():
    is_obj = np.array([0, 0, 1, 0], np.uint8)
    assert oracle.train_early_stop(is_obj, [0, 1, 3], 0, 15, 1)        # no object left (houghforest.rs:303-305)
    assert not oracle.train_early_stop(is_obj, [0, 2], 3, 15, 2)
    assert oracle.train_early_stop(is_obj, [0, 2], 15, 15, 2)          # depth >= max_depth
    assert oracle.train_early_stop(is_obj, [0, 2], 3, 15, 3)           # len < min_subset_size
    hl = train.HoughLearning(10, 80, 80, 15, 1, 100, 0.3, 10, 3, 5.0)
    for idx, depth in (([0, 1, 3], 0), ([0, 2], 3), ([0, 2], 15), ([0, 1, 2], 3)):
        assert hl.early_stop(depth, is_obj[idx]) == oracle.train_early_stop(is_obj, idx, depth, 15, 3)


def test_scale_and_replace_known_answers():
    # the reference's own test vectors (types.rs:454-474) through the vectorised port
    cases = [((100, 100), 0.5, 0.0, 0.0, (0, 0, 50, 50)), ((100, 100), 0.5, 1.0, 1.0, (50, 50, 50, 50)),
             ((100, 100), 0.5, 0.5, 0.5, (25, 25, 50, 50))]
    for (w, h), s, rx, ry, (x, y, ww, hh) in cases:
        x0, y0, x1, y1 = train.scale_and_replace(w, h, s, np.array([rx]), np.array([ry]))
        assert (int(x0[0]), int(y0[0]), int(x1[0] - x0[0]), int(y1[0] - y0[0])) == (x, y, ww, hh)
    # against the oracle's restatement on random arguments, the trained shape included
    rng = np.random.default_rng(1)
    for w, h, s in ((80, 80, 0.3), (64, 48, 0.5), (33, 57, 0.77)):
        u, v = rng.random(50), rng.random(50)
        x0, y0, x1, y1 = train.scale_and_replace(w, h, s, u, v)
        for i in range(50):
            out = np.zeros(4, np.uint32)
            r = np.array([0, 0, w, h], np.uint32)
            oracle.lib().orc_rect_scale_and_replace(r.ctypes.data, s, float(u[i]), float(v[i]), out.ctypes.data)
            assert (int(x0[i]), int(y0[i]), int(x1[i] - x0[i]), int(y1[i] - y0[i])) == tuple(int(t) for t in out)


def test_param_set_shapes():
    hl = train.HoughLearning(10, 80, 80, 15, 1, 100, 0.3, 10, 20, 5.0)
    rects, thr = hl.param_set(np.random.default_rng(0), 500)
    assert rects.shape == (500, 8) and thr.shape == (500,)
    for k in (0, 4):
        assert np.all(rects[:, k + 2] - rects[:, k] == 24) and np.all(rects[:, k + 3] - rects[:, k + 1] == 24)
        assert rects[:, k].min() >= 0 and rects[:, k + 2].max() <= 80
    assert thr.min() >= -256.0 and thr.max() < 256.0
    with pytest.raises(ValueError):
        train.HoughLearning(10, 80, 80, 15, 1, 100, 1.5, 10, 20, 5.0)
    with pytest.raises(ValueError):
        train.HoughLearning(10, 80, 80, 15, 1, 100, 0.3, 0, 20, 5.0)
    with pytest.raises(ValueError):
        train.HoughLearning(10, 80, 80, 15, 1, 100, 0.3, 10, 20, 0.0)


def test_extract_samples_against_oracle_backprojection():
    frames, centres, rots, masks = synth.make_frames(1, seed=3, with_truth=True)
    K = IntrinsicMatrix.default_kinect_intrinsic()
    org, flag, offs, rr = train.extract_samples(frames[0], masks[0], K, centres[0], rots[0], 10, 80, 80)
    assert len(org) == len(flag) == len(offs) == len(rr) > 0 and flag.any() and not flag.all()
    # every window is non-background, in sliding-window order
    for (x0, y0) in org[:50]:
        assert frames[0][y0:y0 + 80, x0:x0 + 80].any()
    assert np.all(np.diff(org[:, 1] * 10000 + org[:, 0]) > 0)
    # offsets = img_to_space_coord(centre pixel) - head centre, in f32 like the reference
    for i in np.flatnonzero(flag)[:20]:
        x, y = org[i] + 40
        p3 = np.zeros(3, np.float32)
        xy = np.array([x, y], np.float32)
        oracle.lib().orc_img_to_space(synth.KINECT_K.ctypes.data, xy.ctypes.data, float(frames[0][y, x]), p3.ctypes.data)
        assert np.array_equal((p3 - centres[0]).view(np.uint32), offs[i].view(np.uint32))
        assert masks[0][y, x] != 0
    assert np.array_equal(rr[0], rots[0].astype(np.float64))

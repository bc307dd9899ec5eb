"""The oracle's restatement of imageproc 0.12's gaussian_blur_f32 (the blur at the end of
build_hough_image, prediction.rs:844) and of predict_parameter_from2dhough (prediction.rs:343-367).
imageproc is an external crate whose source is not under /root/reference: PARITY UNPINNED.  These
tests pin the restatement to the properties its stated construction implies (kernel formula and
radius, unnormalised taps, edge-clamped borders, truncation to u16 after EACH pass, last maximum
wins), so the GPU kernels are compared against a checker whose own behaviour is nailed down."""
import math

import numpy as np
import pytest

import oracle
from depthhead_b200 import synth


@pytest.mark.parametrize("sigma", [0.4, 1.0, 2.5, 8.0])
def test_kernel_construction(sigma):
    k = oracle.gaussian_kernel_f32(sigma)
    r = math.ceil(2.0 * sigma)
    assert len(k) == 2 * r + 1 and np.array_equal(k, k[::-1])
    s32 = np.float32(sigma)
    norm = np.float32(1.0) / (np.sqrt(np.float32(2.0) * np.float32(np.pi)) * s32)
    for i in range(r + 1):
        want = norm * np.exp(-np.float32(i * i) / (np.float32(2.0) * s32 * s32))
        assert abs(float(k[r + i]) - float(want)) <= 2e-7 * float(want) + 1e-12
    assert 0.9 < float(k.sum()) < 1.0 or sigma < 0.6   # the pdf over +-2 sigma: not normalised


def _blur_py(img, sigma):
    """the stated construction in plain numpy (float32, taps in order, clamped borders, truncation per pass)"""
    k = oracle.gaussian_kernel_f32(sigma)
    r = len(k) // 2
    h, w = img.shape

    def one_pass(a, axis):
        n = a.shape[axis]
        acc = np.zeros(a.shape, np.float32)
        for i in range(len(k)):
            idx = np.clip(np.arange(n) + i - r, 0, n - 1)
            acc = acc + np.take(a, idx, axis=axis).astype(np.float32) * k[i]
        out = np.where(acc < np.float32(65535.0), np.where(acc > 0, acc, 0).astype(np.uint16), np.uint16(65535))
        return out.astype(np.uint16)
    return one_pass(one_pass(img, 1), 0)


@pytest.mark.parametrize("shape,sigma", [((20, 30), 2.0), ((7, 5), 3.0), ((1, 9), 1.0), ((33, 2), 0.7), ((64, 48), 8.0)])
def test_blur_equals_the_stated_construction(shape, sigma):
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    for img in (rng.integers(0, 65536, shape).astype(np.uint16), np.full(shape, 65535, np.uint16), np.zeros(shape, np.uint16),
                (rng.random(shape) < 0.05).astype(np.uint16) * 60000):
        assert np.array_equal(oracle.gaussian_blur_u16(img, sigma), _blur_py(img, sigma))


def test_blur_properties():
    # a single bright pixel far from the borders: the response is the truncated outer product of the taps
    a = np.zeros((41, 41), np.uint16)
    a[20, 20] = 50000
    k = oracle.gaussian_kernel_f32(2.0)
    b = oracle.gaussian_blur_u16(a, 2.0)
    row = np.floor(np.float32(50000.0) * k).astype(np.uint16)            # after the horizontal pass
    want_centre_col = np.floor(row[len(k) // 2].astype(np.float32) * k)  # the vertical pass over the centre column
    assert np.array_equal(b[20 - 4:20 + 5, 20], want_centre_col.astype(np.uint16))
    assert np.array_equal(b, b.T) and np.array_equal(b, b[::-1, ::-1])
    # saturation: a full-scale image stays below full scale only because the taps sum to < 1
    full = oracle.gaussian_blur_u16(np.full((9, 9), 65535, np.uint16), 1.0)
    assert full.max() < 65535 and full.min() == full.max()                # clamped borders: constant in, constant out


def test_from2dhough_takes_the_last_maximum_and_backprojects():
    arr = synth.make_forest(seed=3, n_trees=4, max_depth=6)
    of = oracle.OracleForest(arr, 10, 80, 80, 8.0, 20)
    z = np.zeros((480, 640), np.uint16)
    mid, xy = of.predict_parameter_from2dhough(z, synth.KINECT_K)
    assert xy == (639, 479) and np.array_equal(mid, [0.0, 0.0, 0.0])      # all votes 0: max_by_key keeps the last index; z = 0
    d = synth.make_frames(1, seed=11)[0]
    img = of.build_hough_image(d, synth.KINECT_K)
    mid, xy = of.predict_parameter_from2dhough(d, synth.KINECT_K)
    flat = img.reshape(-1)
    best = len(flat) - 1 - int(np.argmax(flat[::-1]))                     # last occurrence of the maximum
    assert xy == (best % 640, best // 640)
    zz = float(d[xy[1], xy[0]])
    assert mid[2] == np.float32(zz)

"""Known-answer tests: the reference's own unit tests, transcribed against the C++ oracle.

Source of every golden value: /root/reference/src/types.rs:454-488 and
/root/reference/src/meancov_estimation.rs:450-533 (tolerances as in the reference).
These pin the oracle's Rect / IntrinsicMatrix / Vec3 / Mat3 / estimate_mean_cov restatement.
"""
import ctypes as C

import numpy as np

import oracle


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _scale(xywh, s, rx, ry):
    out = np.zeros(4, np.uint32)
    oracle.lib().orc_rect_scale_and_replace(_p(np.asarray(xywh, np.uint32)), s, rx, ry, _p(out))
    return out.tolist()


def test_rect_scale_and_replace():  # types.rs:454-474
    r = [1, 2, 10, 20]
    assert _scale(r, 1.0, 0.0, 0.0) == r
    assert _scale(r, 1.0, 0.5, 0.5) == r
    assert _scale(r, 0.5, 0.0, 0.0) == [1, 2, 5, 10]
    assert _scale(r, 0.5, 1.0, 1.0) == [6, 12, 5, 10]
    assert _scale(r, 0.25, 0.5, 0.2) == [4, 5, 2, 5]


def test_intrinsic():  # types.rs:476-488
    L = oracle.lib()
    K = np.array([22.0, 11.4, 12.11, 2.1, 4.1, 2.11, 1.3, 3.1, 19.0], np.float32)
    p3 = np.array([11.0, 12.0, 32.2], np.float32)
    p2 = np.zeros(2, np.float32)
    L.orc_space_to_img(_p(K), _p(p3), _p(p2))
    assert abs(p2[0] - 1.15896578) < 1e-4
    assert abs(p2[1] - 0.21143073) < 1e-4
    back = np.zeros(3, np.float32)
    L.orc_img_to_space(_p(K), _p(p2), float(p3[2]), _p(back))
    assert np.all(np.abs(back - p3) < 1e-4)


def test_kinect_inverse_exact():
    # Kinect K inverse in f32, p3.z == z exactly
    L = oracle.lib()
    K = np.array([560, 0, 320, 0, 560, 240, 0, 0, 1], np.float32)
    inv = np.zeros(9, np.float32)
    L.orc_mat3_inv_f32(_p(K), _p(inv))
    assert inv[8] == np.float32(1.0) and inv[0] == np.float32(560.0) / np.float32(313600.0)
    out = np.zeros(3, np.float32)
    L.orc_img_to_space(_p(K), _p(np.array([100.0, 200.0], np.float32)), 1234.0, _p(out))
    assert out[2] == np.float32(1234.0)


def test_mean_cov2():  # meancov_estimation.rs:450-460
    v = np.array([[2, 6], [3, 4], [3, 8], [4, 6]], np.float64)
    m, c = np.zeros(2), np.zeros(4)
    assert oracle.lib().orc_mean_cov2_f64(_p(v), 4, _p(m), _p(c)) == 0
    assert abs(m[0] - 3) < 1e-3 and abs(m[1] - 6) < 1e-3
    assert abs(c[0] - 0.66666) < 1e-3 and abs(c[3] - 2.6666) < 1e-3
    assert abs(c[1]) < 1e-3 and abs(c[2]) < 1e-3


def test_mean_cov3():  # meancov_estimation.rs:461-490
    L = oracle.lib()
    v = np.array([[1.0, 2.0, 3.0], [1.2, 1.0, 3.2], [-1.0, -2.1, 3.0], [0.0, 1.0, 0.0]], np.float64)
    m, c = np.zeros(3), np.zeros(9)
    assert L.orc_mean_cov3_f64(_p(v), 4, _p(m), _p(c)) == 0
    assert np.allclose(m, [0.3, 0.475, 2.3], atol=1e-3)
    c = c.reshape(3, 3)
    exp = np.array([[1.0266666, 1.576666, 0.36], [1.576666, 3.1691666, -0.49], [0.36, -0.49, 2.36]])
    assert np.all(np.abs(c - exp) < 1e-3)
    v = np.array([[-32.48225021362305, 24.72743034362793, -3.9425208568573],
                  [-25.82341957092285, -25.307233810424805, 1.955498456954956],
                  [35.37421417236328, -18.529083251953125, -5.888242721557617],
                  [43.30265808105469, -60.69481658935547, -15.176074028015137],
                  [32.97354507446289, -7.171285629272461, -3.897606134414673]], np.float64)
    c = np.zeros(9)
    assert L.orc_mean_cov3_f64(_p(v), 5, _p(m), _p(c)) == 0
    exp = np.array([[1341.63076476, -685.47821414, -157.223241746],
                    [-685.478214144, 954.396794746, 110.60252659],
                    [-157.223241746, 110.60252659, 38.573568346]])
    assert np.all(np.abs(c.reshape(3, 3) - exp) < 1e-3)
    assert abs(L.orc_mat3_det_f64(_p(c)) - 15102509.494226849) < 1e-4


def test_mean_cov_single_sample_is_nan():  # n == 1 -> 0/0 -> NaN -> gate false
    L = oracle.lib()
    v = np.array([[1.0, 2.0, 3.0]], np.float64)
    m, c = np.zeros(3), np.zeros(9)
    L.orc_mean_cov3_f64(_p(v), 1, _p(m), _p(c))
    assert np.isnan(L.orc_mat3_trace_f64(_p(c)))
    vf = v.astype(np.float32)
    mf, cf = np.zeros(3, np.float32), np.zeros(9, np.float32)
    L.orc_mean_cov3_f32(_p(vf), 1, _p(mf), _p(cf))
    assert np.isnan(cf[0] + cf[4] + cf[8])


def test_det_2_3_trace():  # meancov_estimation.rs:492-500
    L = oracle.lib()
    m2 = np.array([1.0, 3.0, 2.0, 44.0])
    assert abs(L.orc_mat2_det_f64(_p(m2)) - 38.0) < 1e-3
    assert abs(L.orc_mat2_trace_f64(_p(m2)) - 45.0) < 1e-3
    m3 = np.array([1.0, 3.0, 22.0, 2.0, 44.0, 1.0, 2.0, 0.0, 3.1])
    assert abs(L.orc_mat3_det_f64(_p(m3)) - -1812.199) < 1e-3
    assert abs(L.orc_mat3_trace_f64(_p(m3)) - 48.1) < 1e-3


def test_inverse():  # meancov_estimation.rs:502-515
    L = oracle.lib()
    m2 = np.array([4.3, 2.4, 2.1, 424.11])
    assert abs(L.orc_mat2_det_f64(_p(m2)) - 1818.633) < 1e-3
    o2 = np.zeros(4)
    L.orc_mat2_inv_f64(_p(m2), _p(o2))
    assert np.all(np.abs(o2 - [0.23320263, -0.00131967, -0.00115471, 0.00236441]) < 1e-3)
    m3 = np.array([2.3, 1.4, 12.11, 2.1, 44.11, 2.11, 1.3, 4.1, 19.0])
    o3 = np.zeros(9)
    L.orc_mat3_inv_f64(_p(m3), _p(o3))
    exp = [0.65540671, 0.01821446, -0.4197583, -0.02936075, 0.02209108, 0.01626034, -0.03850788, -0.00601328,
           0.07784307]
    assert np.all(np.abs(o3 - exp) < 1e-3)


def test_mat_vec_mul():  # meancov_estimation.rs:517-525
    L = oracle.lib()
    o2 = np.zeros(2)
    L.orc_mat2_mul_vec2_f64(_p(np.array([1.3, 12.1, 3.1, 33.1])), _p(np.array([11.0, 12.0])), _p(o2))
    assert np.all(np.abs(o2 - [159.5, 431.3]) < 1e-3)
    o3 = np.zeros(3)
    L.orc_mat3_mul_vec3_f64(_p(np.array([1.3, 12.1, 2.3, 3.1, 33.1, 14.1, 1.0, 2.0, 3.0])),
                            _p(np.array([11.0, 12.0, 32.2])), _p(o3))
    assert np.all(np.abs(o3 - [233.56, 885.32, 131.6]) < 1e-3)


def test_transposed_matrix_via_cov():  # meancov_estimation.rs:527-533 (outer product v v^T)
    # cov of {v, -v} about mean 0 with n-1 == 1 is 2 * v v^T
    L = oracle.lib()
    v = np.array([[2.0, 1.1, 4.3], [-2.0, -1.1, -4.3]])
    m, c = np.zeros(3), np.zeros(9)
    L.orc_mean_cov3_f64(_p(v), 2, _p(m), _p(c))
    exp = 2 * np.array([[4.0, 2.2, 8.6], [2.2, 1.21, 4.73], [8.6, 4.73, 18.49]])
    assert np.all(np.abs(c.reshape(3, 3) - exp) < 1e-3)


def test_kernel_table():  # meanshift.rs:228-252; sigma used as the variance (prediction.rs:314)
    k = np.zeros(8000, np.float32)
    oracle.lib().orc_build_kernel(20, 8.0, _p(k))
    k3 = k.reshape(20, 20, 20)  # [z][y][x]
    assert k3[10, 10, 10] == np.float32(1.0)
    assert k3[10, 10, 11] == np.exp(np.float32(-1.0) / np.float32(16.0), dtype=np.float32)
    assert k3[0, 0, 0] == np.exp(np.float32(-300.0) / np.float32(16.0), dtype=np.float32)
    assert np.array_equal(k3, k3.transpose(2, 1, 0))

"""A SECOND, independent restatement of the prediction path, for small cases only.

TEST INFRASTRUCTURE (like everything under oracle/): only tests/ may import it.  It exists to take
the single-reader risk out of the C++ oracle: this file was written directly from the Rust source
(src/hough/prediction.rs:421-753, src/meanshift.rs:228-252,328-407,
src/meancov_estimation.rs:201-216,260-265,290-304,339-378, src/types.rs:317-339,424-445,
src/hough/houghforest.rs:185-193), not from oracle/depthhead_oracle.cpp, in a different language
and with different data structures (dicts keyed by integer triples, numpy scalars of the
reference's widths, pixel loops instead of tables).  tests/test_second_opinion.py requires both
restatements to agree bit for bit on every intermediate they share.  Agreement of two readers is
still not the Rust binary: the parts that live in stamm 0.2.0 (tree storage, which child
`Binar::One` takes) stay PARITY UNPINNED in both.

Pure-Python loops: a 200 x 160 frame with a 4-tree forest takes a few seconds.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
f64 = np.float64

GUESS_GRID_PARTS = 20      # prediction.rs:270-286
ROT_GRID_PARTS = 120
MAX_VARIANCE_ROT = 400.0
MAX_VARIANCE_OFFSET = f32(5200.0)
PI_REF = 3.14159           # the reference's own constant (prediction.rs:448-450, 477-482)


def as_i32(x) -> int:
    """Rust `as i32` on a float: toward zero, saturating, NaN -> 0"""
    x = float(x)
    if math.isnan(x):
        return 0
    if x >= 2147483647.0:
        return 2147483647
    if x <= -2147483648.0:
        return -2147483648
    return int(x)


def as_usize(x) -> int:
    x = float(x)
    if math.isnan(x) or x <= 0.0:
        return 0
    if x >= 1.8446744073709552e19:
        return 2 ** 64 - 1
    return int(x)


def wrap_u32(x: int) -> int:
    return x & 0xFFFFFFFF


# ---------------------------------------------------------------- meancov_estimation.rs
def mat3_mul_vec3_f32(m, v):
    """Mul<Vec3> for Mat3 (:201-216): tmp = v0*m[j][0]; tmp = tmp + v_i*m[j][i]"""
    out = []
    for j in range(3):
        tmp = f32(v[0]) * f32(m[j][0])
        for i in (1, 2):
            tmp = f32(tmp + f32(f32(v[i]) * f32(m[j][i])))
        out.append(f32(tmp))
    return out


def mat3_inv_f32(m):
    """Mat3::inv (:344-352): adjugate / det, all f32"""
    a, b, c = (f32(x) for x in m[0])
    d, e, f = (f32(x) for x in m[1])
    g, h, i = (f32(x) for x in m[2])
    det = f32(f32(f32(a * f32(f32(e * i) - f32(f * h))) - f32(d * f32(f32(b * i) - f32(c * h)))) + f32(g * f32(f32(b * f) - f32(c * e))))
    adj = [[f32(e * i) - f32(f * h), f32(c * h) - f32(b * i), f32(b * f) - f32(c * e)],
           [f32(f * g) - f32(d * i), f32(a * i) - f32(c * g), f32(c * d) - f32(a * f)],
           [f32(d * h) - f32(e * g), f32(b * g) - f32(a * h), f32(a * e) - f32(b * d)]]
    with np.errstate(all="ignore"):
        return [[f32(f32(x) / det) for x in row] for row in adj]


def trace_of_cov(votes, dtype):
    """estimate_mean_cov(set).1.trace() (:359-378, :260-265).  Only the diagonal of the covariance
    matters for the trace; rotations in f64, offsets in f32 with the divisors cast f64 -> f32
    (:290-304)."""
    n = len(votes)
    t = dtype
    mean = [t(votes[0][k]) for k in range(3)]
    for i in range(1, n):
        mean = [t(mean[k] + t(votes[i][k])) for k in range(3)]
    with np.errstate(all="ignore"):
        div = t(f64(n))
        mean = [t(mean[k] / div) for k in range(3)]
        diag = []
        for k in range(3):
            d0 = t(t(votes[0][k]) - mean[k])
            diag.append(t(d0 * d0))
        for i in range(1, n):
            for k in range(3):
                dk = t(t(votes[i][k]) - mean[k])
                diag[k] = t(diag[k] + t(dk * dk))
        div1 = t(f64(n - 1))
        tr = t(0.0)
        for k in range(3):
            tr = t(tr + t(diag[k] / div1))
    return tr


# ---------------------------------------------------------------- types.rs
class Intrinsic:
    def __init__(self, K):
        self.k = [[f32(K[j][i]) for i in range(3)] for j in range(3)]
        self.inv = mat3_inv_f32(self.k)

    def img_to_space(self, x, y, z):  # :432-445
        res = mat3_mul_vec3_f32(self.inv, [f32(x), f32(y), f32(1.0)])
        with np.errstate(all="ignore"):
            c = f32(f32(z) / res[2])
            return [f32(r * c) for r in res]

    def space_to_img(self, p):  # :424-428
        res = mat3_mul_vec3_f32(self.k, p)
        with np.errstate(all="ignore"):
            return [f32(res[0] / res[2]), f32(res[1] / res[2])]


def average_value_in_rect(img, ox, oy, rect):
    """SubImage::average_value_in_rect (types.rs:317-339): rect = (x0, y0, x1, y1) relative to the
    sub-image origin (ox, oy); u64 sum, f64 division, 0.0 for an empty rectangle"""
    x0, y0, x1, y1 = rect
    total, count = 0, 0
    for y in range(oy + y0, oy + y1):
        for x in range(ox + x0, ox + x1):
            count += 1
            total += int(img[y, x])
    if count == 0:
        return f64(0.0)
    return f64(total) / f64(count)


# ---------------------------------------------------------------- the forest (flat arrays of synth.make_forest)
def walk(arr, t, img, ox, oy):
    """root -> leaf with HoughTreeFunctions::binarize (houghforest.rs:185-193); the container
    semantics are this repository's (child[bit], bit = avg1 - avg2 > threshold; < 0 = ~leaf)"""
    n0, l0 = int(arr["tree_node_off"][t]), int(arr["tree_leaf_off"][t])
    if int(arr["tree_node_off"][t + 1]) == n0:
        return l0
    node = 0
    while True:
        r = arr["rects"][n0 + node]
        avg1 = average_value_in_rect(img, ox, oy, (int(r[0]), int(r[1]), int(r[2]), int(r[3])))
        avg2 = average_value_in_rect(img, ox, oy, (int(r[4]), int(r[5]), int(r[6]), int(r[7])))
        bit = 1 if (avg1 - avg2) > f64(arr["threshold"][n0 + node]) else 0
        ch = int(arr["child"][n0 + node][bit])
        if ch < 0:
            return l0 + (~ch)
        node = ch


def kernel_value(dx, dy, dz, sigma):
    """kernel_function (meanshift.rs:228-232) through libm's expf, as Rust's f32::exp does"""
    import ctypes
    import ctypes.util
    global _LIBM
    try:
        _LIBM
    except NameError:
        _LIBM = ctypes.CDLL(ctypes.util.find_library("m"))
        _LIBM.expf.restype = ctypes.c_float
        _LIBM.expf.argtypes = [ctypes.c_float]
    norm = dx * dx + dy * dy + dz * dz
    arg = f32(f32(f32(-1.0) * f32(norm)) / f32(f32(2.0) * f32(sigma)))
    return f32(_LIBM.expf(float(arg)))


def meanshift(acc: dict, init, sigma, iterations):
    """MeanShift::meanshift (meanshift.rs:328-407) on a dict-backed SparseArray3D<u32>"""
    pos = [int(init[0]), int(init[1]), int(init[2])]
    trace = []
    zero = False
    for _ in range(iterations):
        num = [f32(0.0), f32(0.0), f32(0.0)]
        den = f32(0.0)
        for x in range(-10, 10):
            for y in range(-10, 10):
                for z in range(-10, 10):
                    factor = acc.get((pos[0] + x, pos[1] + y, pos[2] + z), 0)
                    if factor == 0:
                        continue
                    influence = kernel_value(x, y, z, sigma)
                    w = f32(influence * f32(factor))
                    ap = [f32(pos[0] + x), f32(pos[1] + y), f32(pos[2] + z)]
                    num = [f32(num[k] + f32(ap[k] * w)) for k in range(3)]
                    den = f32(den + w)
        if den == f32(0.0):
            zero = True
            break
        with np.errstate(all="ignore"):
            pos = [as_i32(f32(num[k] / den)) for k in range(3)]
        trace.append(tuple(pos))
    return tuple(pos), trace, zero


def predict(arr, depth, K, stepwidth, sub_w, sub_h, sigma, iterations, midp_guess=None, rot_guess=None):
    """predict_parameter_generic (prediction.rs:421-493) with build_hough_cube_generic (:509-753).
    Returns a dict of every intermediate."""
    img = np.asarray(depth)
    h, w = img.shape
    intr = Intrinsic(K)
    T = int(arr["n_trees"])
    guess_pos = [0] * (GUESS_GRID_PARTS * GUESS_GRID_PARTS)
    guess_rot = {}                                  # (x, y, z) -> u32, dense 20^3 in the reference
    mid, rot = {}, {}
    left_w, left_h = sub_w // 2, sub_h // 2
    right_w, right_h = sub_w - left_w, sub_h - left_h
    leaf_rows, gate_rows, p3_rows = [], [], []
    static = {}
    y = left_h
    while y < h - right_h:
        x = left_w
        while x < w - right_w:
            z = int(img[y, x])
            p3 = intr.img_to_space(x, y, z)
            ox, oy = x - left_w, y - left_h
            if average_value_in_rect(img, ox, oy, (0, 0, sub_w, sub_h)) > 0.0:
                leafs = [walk(arr, t, img, ox, oy) for t in range(T)]
                s = f64(0.0)
                for L in leafs:
                    s = f64(s + f64(arr["prob"][L]))
                prob = f64(s / f64(len(leafs)))
                gate = bool(prob > 0.7)
                if gate:
                    for L in leafs:
                        lp = f64(arr["prob"][L])
                        if not (lp > 0.0):
                            continue
                        v0, v1 = int(arr["vote_off"][L]), int(arr["vote_off"][L + 1])
                        if L not in static:
                            offs = [arr["offsets"][i] for i in range(v0, v1)]
                            rots = [arr["rotations"][i] for i in range(v0, v1)]
                            valtoadd = wrap_u32(as_usize(f64(1000.0) * lp) // (v1 - v0))
                            static[L] = (valtoadd, bool(trace_of_cov(rots, f64) <= MAX_VARIANCE_ROT),
                                         bool(trace_of_cov(offs, f32) <= MAX_VARIANCE_OFFSET))
                        valtoadd, rot_ok, off_ok = static[L]
                        if rot_ok:
                            for i in range(v0, v1):
                                rr = []
                                for k in range(3):
                                    r = as_i32(f64(f64(arr["rotations"][i][k]) * f64(ROT_GRID_PARTS)) / f64(360.0)) + ROT_GRID_PARTS // 2
                                    if r >= ROT_GRID_PARTS:
                                        r -= ROT_GRID_PARTS
                                    elif r < 0:
                                        r += ROT_GRID_PARTS
                                    rr.append(r)
                                rough = tuple(wrap_u32(r) * GUESS_GRID_PARTS // ROT_GRID_PARTS for r in rr)
                                rot[tuple(rr)] = wrap_u32(rot.get(tuple(rr), 0) + valtoadd)
                                guess_rot[rough] = wrap_u32(guess_rot.get(rough, 0) + valtoadd)
                        if off_ok:
                            for i in range(v0, v1):
                                np_ = [f32(p3[k] - f32(arr["offsets"][i][k])) for k in range(3)]
                                if np_[2] < 0.0:
                                    continue
                                p2 = intr.space_to_img(np_)
                                # max!/min! (prediction.rs:19-25): `if a > b {a} else {b}` / `if a < b {a} else {b}`
                                mx = p2[0] if p2[0] > f32(0.0) else f32(0.0)
                                x2d = mx if mx < f32(w - 1) else f32(w - 1)
                                my = p2[1] if p2[1] > f32(0.0) else f32(0.0)
                                y2d = my if my < f32(h - 1) else f32(h - 1)
                                key = (as_i32(np_[0]), as_i32(np_[1]), as_i32(np_[2]))
                                mid[key] = wrap_u32(mid.get(key, 0) + valtoadd)
                                gx = as_usize(x2d) * GUESS_GRID_PARTS // w
                                gy = as_usize(y2d) * GUESS_GRID_PARTS // h
                                guess_pos[gy * GUESS_GRID_PARTS + gx] = wrap_u32(guess_pos[gy * GUESS_GRID_PARTS + gx] + valtoadd)
                leaf_rows.append(leafs)
                gate_rows.append(gate)
            else:
                leaf_rows.append([-1] * T)
                gate_rows.append(False)
            p3_rows.append(p3)
            x += stepwidth
        y += stepwidth
    # ---- seeds (:694-752)
    prev_max, best_idx = 0, 0
    for idx, el in enumerate(guess_pos):
        if el > prev_max:
            prev_max, best_idx = el, idx
    gw, gh = w // GUESS_GRID_PARTS, h // GUESS_GRID_PARTS
    mxg, myg = best_idx % GUESS_GRID_PARTS, best_idx // GUESS_GRID_PARTS
    zs, zn = 0, 0
    for yy in range(gh * myg, gh * myg + gh):
        for xx in range(gw * mxg, gw * mxg + gw):
            if int(img[yy, xx]) > 0:
                zs += int(img[yy, xx])
                zn += 1
    meanz = f32(f64(zs) / f64(zn)) if zn > 0 else f32(0.0)
    max_x = f32(f32(f32(mxg) + f32(0.5)) * f32(gw))
    max_y = f32(f32(f32(myg) + f32(0.5)) * f32(gh))
    max3d = intr.img_to_space(max_x, max_y, meanz)
    guessmid = (as_i32(max3d[0]), as_i32(max3d[1]), as_i32(max3d[2]))
    best = (0, 0, 0, 0)
    for zc in range(GUESS_GRID_PARTS):          # FullArray3DIter: x fastest, then y, then z
        for yc in range(GUESS_GRID_PARTS):
            for xc in range(GUESS_GRID_PARTS):
                c = guess_rot.get((xc, yc, zc), 0)
                if c > 0 and c > best[3]:
                    best = (xc, yc, zc, c)
    guessrot = tuple(f64(f64(f64(best[k]) * f64(360.0)) + f64(180.0)) / f64(GUESS_GRID_PARTS) for k in range(3))
    # ---- predict_parameter_generic (:421-493)
    if midp_guess is not None:
        guessmid = tuple(as_i32(f32(g)) for g in midp_guess)
    if rot_guess is not None:
        guessrot = tuple(f64(f64(f64(g) * f64(180.0)) / f64(PI_REF)) + f64(180.0) for g in rot_guess)
    seed_rot = tuple(as_i32(f64(f64(g) * f64(ROT_GRID_PARTS)) / f64(360.0)) for g in guessrot)
    res_mid, tr_mid, zero_mid = meanshift(mid, guessmid, sigma, iterations)
    res_rot, tr_rot, zero_rot = meanshift(rot, seed_rot, sigma, iterations)
    rotation = [f64(f64(f64(r) - f64(ROT_GRID_PARTS) / f64(2.0)) / f64(ROT_GRID_PARTS // 2)) * f64(PI_REF) for r in res_rot]
    return {"leaf": np.array(leaf_rows, np.int32), "gate": np.array(gate_rows, bool), "p3": np.array(p3_rows, np.float32),
            "guess_pos": np.array(guess_pos, np.uint32), "guess_rot": guess_rot, "mid": mid, "rot": rot,
            "seed_mid": guessmid, "seed_rot": seed_rot, "ms_mid": tr_mid, "ms_rot": tr_rot,
            "mid_point": np.array([f32(v) for v in res_mid], np.float32), "rotation": np.array(rotation, np.float64)}

"""Biwi wire formats on the CPU: the oracle's restatement of
src/db_reader/biwi.rs against hand-built byte streams, the synthetic encoder against the oracle,
and the product's host-only parsers (depth header, depth.cal, pose file) against the oracle.
The run-length expansion itself runs on the GPU: tests/test_gpu_biwi.py."""
import struct

import numpy as np
import pytest

import oracle
from depthhead_b200 import DhError, IntrinsicMatrix, biwi, synth

KINECT = IntrinsicMatrix.default_kinect_intrinsic()


def _stream(w, h, runs):
    out = [struct.pack("<II", w, h)]
    for n_empty, pixels in runs:
        out.append(struct.pack("<II", n_empty, len(pixels)))
        out.append(np.asarray(pixels, "<u2").tobytes())
    return b"".join(out)


def test_read_depth_known_streams():
    # biwi.rs:81-103 by hand: 4x2 frame = [0 0 7 8 | 9 0 0 5]
    s = _stream(4, 2, [(2, [7, 8, 9]), (2, [5])])
    assert np.array_equal(oracle.biwi_read_depth(s), np.array([[0, 0, 7, 8], [9, 0, 0, 5]], np.uint16))
    # leading pixels: an empty run of length 0; trailing zeros: a full run of length 0
    s = _stream(3, 1, [(0, [4]), (2, [])])
    assert np.array_equal(oracle.biwi_read_depth(s), np.array([[4, 0, 0]], np.uint16))
    # bytes after the last pixel are never read (biwi.rs:89)
    assert np.array_equal(oracle.biwi_read_depth(s + b"garbage"), np.array([[4, 0, 0]], np.uint16))
    # 0x0 image: no run is read at all
    assert oracle.biwi_read_depth(struct.pack("<II", 0, 0)).size == 0
    # little endian u16
    s = struct.pack("<II", 1, 1) + struct.pack("<II", 0, 1) + bytes([0x34, 0x12])
    assert oracle.biwi_read_depth(s)[0, 0] == 0x1234


def test_read_depth_failures_like_the_reference():
    ok = _stream(4, 1, [(1, [1, 2, 3])])
    for cut in (3, 7, 9, 15, len(ok) - 1):              # UnexpectedEof at every read site
        with pytest.raises(oracle.BiwiError) as e:
            oracle.biwi_read_depth(ok[:cut])
        assert e.value.code == 1
    with pytest.raises(oracle.BiwiError) as e:          # full run past the last pixel: unwrap on None (:97)
        oracle.biwi_read_depth(_stream(2, 1, [(1, [5, 6])]))
    assert e.value.code == 2
    with pytest.raises(oracle.BiwiError) as e:          # empty run past the last pixel (:92)
        oracle.biwi_read_depth(_stream(2, 1, [(3, [])]))
    assert e.value.code == 2
    with pytest.raises(oracle.BiwiError) as e:          # runs of length 0 forever: reads until the file ends
        oracle.biwi_read_depth(_stream(2, 1, [(0, []), (0, []), (0, [])]))
    assert e.value.code == 1


@pytest.mark.parametrize("shape,seed", [((480, 640), 1), ((7, 13), 2), ((1, 1), 3), ((33, 250), 4)])
def test_encoder_round_trip(shape, seed):
    rng = np.random.default_rng(seed)
    if shape == (480, 640):
        frames = list(synth.make_frames(2, seed=seed))
    else:
        frames = [np.where(rng.random(shape) < 0.4, rng.integers(1, 65536, shape), 0).astype(np.uint16) for _ in range(3)]
    frames += [np.zeros(shape, np.uint16), np.full(shape, 65535, np.uint16)]
    for f in frames:
        data = biwi.encode_depth(f)
        assert biwi.depth_dims(data) == (shape[1], shape[0])
        assert np.array_equal(oracle.biwi_read_depth(data), f)


def test_pack_files_layout():
    files = [biwi.encode_depth(np.full((2, 3), v, np.uint16)) for v in (0, 5, 9)]
    blob, off = biwi.pack_files(files)
    assert off[0] == 0 and all(int(o) % 16 == 0 for o in off) and len(off) == 4
    for i, f in enumerate(files):
        assert blob[int(off[i]):int(off[i]) + len(f)].tobytes() == f


CAL = "575.816 0 320 \n0 575.816 240 \n0 0 1 \n\n0 0 0 0 \n\n1 0 0 \n0 1 0 \n0 0 1 \n\n0 0 0 \n\n640 480\n"


def test_read_cal_matches_oracle():
    K = oracle.biwi_read_cal(CAL)
    assert np.array_equal(K, np.array([[575.816, 0, 320], [0, 575.816, 240], [0, 0, 1]], np.float32))
    assert np.array_equal(biwi.read_cal(CAL).mat, K)
    got = np.zeros(9, np.float32)
    from depthhead_b200 import capi
    capi.check(capi.load().dh_biwi_parse_cal(CAL.encode(), len(CAL), capi.ptr(got)))
    assert np.array_equal(got.reshape(3, 3), K)
    # the regex (\d+[\.\d+]*) never sees a sign: -1.5 reads as 1.5 (biwi.rs:30)
    t = "-1.5 2 3\n4 -5 6\n7 8 -9e1\n"   # "9e1" -> tokens "9" and "1": four numbers on the line
    with pytest.raises(oracle.BiwiError) as e:
        oracle.biwi_read_cal(t)
    assert e.value.code == 3
    with pytest.raises(DhError):
        biwi.read_cal(t)
    t = "-1.5 2 3\n4 -5 6\n7 8 -9\n"
    K = oracle.biwi_read_cal(t)
    assert np.array_equal(K, np.array([[1.5, 2, 3], [4, 5, 6], [7, 8, 9]], np.float32))
    capi.check(capi.load().dh_biwi_parse_cal(t.encode(), len(t), capi.ptr(got)))
    assert np.array_equal(got.reshape(3, 3), K)
    for bad, code in (("1 2\n3 4 5\n6 7 8\n", 1), ("1 2 3\n4 5 6\n", 1), ("1.2.3 4 5\n1 2 3\n1 2 3\n", 2), ("1+2 4 5\n1 2 3\n1 2 3\n", 2),
                      ("1 2 3 4\n1 2 3\n1 2 3\n", 3), ("", 1)):
        with pytest.raises(oracle.BiwiError) as e:
            oracle.biwi_read_cal(bad)
        assert e.value.code == code, bad
        with pytest.raises(DhError):
            biwi.read_cal(bad)
    # f32::from_str is correctly rounded; so is strtof
    t = "0.1 16777217 3.4028235e3\n1 2 3\n1 2 3\n".replace("e3", "")
    K = oracle.biwi_read_cal(t)
    assert K[0, 0] == np.float32(0.1) and K[0, 1] == np.float32(16777216.0)


def test_read_gt_matches_oracle():
    vals = [12.5, -40.25, 880.0, 0.1, -0.2, 0.3]
    data = struct.pack("<6f", *vals)
    p3, p2, rot = oracle.biwi_read_gt(data, synth.KINECT_K)
    assert np.array_equal(p3, np.float32(vals[:3])) and np.array_equal(rot, np.float32(vals[3:]))
    # types.rs:424-428 with the Kinect matrix: x*560/z + 320 in f32, products and sums unfused
    ex = np.float32(np.float32(np.float32(12.5) * np.float32(560)) + np.float32(np.float32(880) * np.float32(320))) / np.float32(880)
    assert p2[0] == ex
    got = biwi.read_gt(data + b"trailing", KINECT)
    assert np.array_equal(got["pos3d"], p3) and np.array_equal(got["pos2d"].view(np.uint32), p2.view(np.uint32))
    assert np.array_equal(got["rot"], rot)
    with pytest.raises(oracle.BiwiError):
        oracle.biwi_read_gt(data[:23], synth.KINECT_K)
    with pytest.raises(DhError):
        biwi.read_gt(data[:23], KINECT)


def test_depth_dims_errors():
    with pytest.raises(DhError):
        biwi.depth_dims(b"\x01\x00\x00")


# ---- the library's own writer (dh_biwi_encode_depth): what dh_predict_batch runs on its worker
#      threads for host frames; checked against the oracle's restatement of the reference reader
def _edge_frames(h, w, seed):
    rng = np.random.default_rng(seed)
    fr = []
    fr.append(np.zeros((h, w), np.uint16))                                   # all background
    fr.append(rng.integers(1, 65536, (h, w)).astype(np.uint16))              # no background at all
    a = np.zeros((h, w), np.uint16); a.flat[0] = 7; fr.append(a)             # first pixel only
    a = np.zeros((h, w), np.uint16); a.flat[-1] = 9; fr.append(a)            # last pixel only (inside the partial tail group)
    a = rng.integers(1, 4000, (h, w)).astype(np.uint16); a[rng.random((h, w)) < 0.5] = 0; fr.append(a)   # salt and pepper
    a = np.zeros((h, w), np.uint16); a.flat[::16] = 1; fr.append(a)          # one pixel in every 16-pixel group
    a = np.zeros((h, w), np.uint16); a.flat[15::32] = 1; fr.append(a)        # every other group
    a = rng.integers(1, 4000, (h, w)).astype(np.uint16); a.flat[: (h * w) // 2] = 0; fr.append(a)        # one long run each
    return np.stack(fr)


@pytest.mark.parametrize("shape", [(480, 640), (7, 13), (1, 1), (33, 250), (1, 16), (3, 16), (2, 17), (5, 31)])
@pytest.mark.parametrize("threads", [1, 3])
def test_library_encoder_round_trips_through_the_reference_reader(shape, threads):
    h, w = shape
    frames = _edge_frames(h, w, seed=h * 1000 + w)
    if shape == (480, 640):
        frames = np.concatenate([frames, synth.make_frames(3, seed=5)])
    blob, offsets = biwi.encode_frames(frames, threads=threads)
    assert len(offsets) == len(frames) + 1 and all(int(o) % 16 == 0 for o in offsets)
    L = __import__("depthhead_b200").capi.load()
    bound = int(L.dh_biwi_encode_bound(w, h))
    for i, f in enumerate(frames):
        data = blob[int(offsets[i]):int(offsets[i + 1])].tobytes()
        assert len(data) <= ((bound + 15) & ~15)
        assert biwi.depth_dims(data) == (w, h)
        assert np.array_equal(oracle.biwi_read_depth(data), f), (shape, i)
    # mostly-background frames shrink
    if shape == (480, 640):
        assert int(offsets[1]) - int(offsets[0]) < 64                       # all background: header + one run
        syn = int(offsets[len(frames)]) - int(offsets[len(frames) - 3])
        assert syn < 3 * h * w * 2 / 3


def test_library_encoder_empty_batch_and_sizing():
    blob, offsets = biwi.encode_frames(np.zeros((0, 4, 4), np.uint16))
    assert len(offsets) == 1 and offsets[0] == 0
    # a too-small buffer reports the size it needs and writes nothing past cap
    import ctypes as C
    from depthhead_b200 import capi
    fr = np.ones((2, 8, 8), np.uint16)
    need = C.c_size_t(0)
    off = np.zeros(3, np.uint64)
    small = np.full(32, 0xEE, np.uint8)
    capi.check(capi.load().dh_biwi_encode_depth(capi.ptr(fr), 2, 8, 8, 1, capi.ptr(small), 16, capi.ptr(off), C.byref(need)))
    assert need.value >= 2 * (8 + 8 + 128) and np.all(small[16:] == 0xEE)

"""GPU side of the Biwi wire formats: the run-length expansion of
read_depth (src/db_reader/biwi.rs:81-103) in biwi_decode_kernel against the oracle, the failures
the reference reader has, and read_depth + predict straight from compressed files."""
import struct

import numpy as np
import pytest

import oracle
from depthhead_b200 import Context, DhError, HoughPrediction, IntrinsicMatrix, biwi, capi, synth

pytestmark = pytest.mark.gpu

K = IntrinsicMatrix.default_kinect_intrinsic()


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c
    c.close()


def _stream(w, h, runs):
    out = [struct.pack("<II", w, h)]
    for n_empty, pixels in runs:
        out.append(struct.pack("<II", n_empty, len(pixels)))
        out.append(np.asarray(pixels, "<u2").tobytes())
    return b"".join(out)


def test_decode_matches_oracle_on_synthetic_frames(ctx):
    frames = synth.make_frames(6, seed=9)
    files = [biwi.encode_depth(f) for f in frames]
    got = biwi.read_depth(files, ctx=ctx)
    assert got.shape == frames.shape
    for i, f in enumerate(files):
        assert np.array_equal(got[i], oracle.biwi_read_depth(f)), "frame %d" % i
    assert np.array_equal(got, frames)
    assert np.array_equal(biwi.read_depth(files[0], ctx=ctx), frames[0])


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (7, 13), (33, 250), (120, 97), (480, 640)])
def test_decode_edge_frames(ctx, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    h, w = shape
    cases = [
        np.zeros(shape, np.uint16),                                   # one empty run
        np.full(shape, 65535, np.uint16),                             # one full run longer than every segment / piece
        (np.indices(shape).sum(0) % 2 * 777).astype(np.uint16),       # a run per pixel: 10 bytes of file per pixel
        np.where(rng.random(shape) < 0.5, rng.integers(1, 65536, shape), 0).astype(np.uint16),
        np.where(rng.random(shape) < 0.02, rng.integers(1, 65536, shape), 0).astype(np.uint16),   # long empty runs
        np.where(rng.random(shape) < 0.98, rng.integers(1, 65536, shape), 0).astype(np.uint16),   # long full runs
    ]
    one = np.zeros(shape, np.uint16)
    one[h - 1, w - 1] = 9                                             # the very last pixel
    cases.append(one)
    files = [biwi.encode_depth(c) for c in cases]
    got = biwi.read_depth(files, ctx=ctx)
    for i, c in enumerate(cases):
        assert np.array_equal(oracle.biwi_read_depth(files[i]), c)
        assert np.array_equal(got[i], c), "case %d" % i


def test_decode_hand_built_streams(ctx):
    # runs of length zero, pixels first, data after the last pixel, headers on odd 2-byte positions
    s1 = _stream(4, 2, [(2, [7, 8, 9]), (2, [5])])
    s2 = _stream(4, 2, [(0, [4]), (0, []), (0, [1, 2]), (5, [])]) + b"tail that is never read"
    s3 = _stream(4, 2, [(0, []), (8, [])])
    got = biwi.read_depth([s1, s2, s3], ctx=ctx)
    for g, s in zip(got, (s1, s2, s3)):
        assert np.array_equal(g, oracle.biwi_read_depth(s))


def test_decode_failures_like_the_reference(ctx):
    good = biwi.encode_depth(synth.make_frames(1, seed=4)[0])
    bad_cases = [
        good[: len(good) // 2],                    # UnexpectedEof inside a full run
        good[:12],                                 # inside a run header
        _stream(640, 480, [(640 * 480 + 1, [])]),  # empty run past the last pixel (panic)
        _stream(640, 480, [(640 * 480 - 1, [1, 2])]),
        _stream(640, 480, [(0, [])] * 50),         # zero-length runs until the file ends
        _stream(320, 240, [(320 * 240, [])]),      # valid file of another size
    ]
    for i, bad in enumerate(bad_cases):
        if i < 5:
            with pytest.raises(oracle.BiwiError):
                oracle.biwi_read_depth(bad)
        blob, off = biwi.pack_files([good, bad, good])
        out = np.zeros((3, 480, 640), np.uint16)
        rc = capi.load().dh_biwi_decode_depth(ctx._h, capi.ptr(blob), capi.ptr(off), 3, 640, 480, capi.ptr(out), capi.DH_DEPTH_HOST)
        assert rc == capi.DH_E_ARG, "case %d" % i
        assert b"frame 1" in capi.load().dh_last_error()
    # misaligned / non-monotone offsets are argument errors
    blob, off = biwi.pack_files([good, good])
    off2 = off.copy()
    off2[1] += 2
    out = np.zeros((2, 480, 640), np.uint16)
    assert capi.load().dh_biwi_decode_depth(ctx._h, capi.ptr(blob), capi.ptr(off2), 2, 640, 480, capi.ptr(out), 0) == capi.DH_E_ARG
    # the context still works afterwards
    assert np.array_equal(biwi.read_depth(good, ctx=ctx), oracle.biwi_read_depth(good))


def test_predict_from_compressed_files(ctx):
    arr = synth.make_forest(seed=3, n_trees=4, max_depth=7)
    js = synth.forest_to_json(arr, stepwidth=8)
    hp = HoughPrediction.from_json(js)
    of = oracle.OracleForest.from_json(js)
    frames = synth.make_frames(11, seed=31)
    files = [biwi.encode_depth(f) for f in frames]
    blob, off = biwi.pack_files(files)
    ctx.set_chunk_frames(4)   # several chunks over both lanes and both staging slots
    try:
        got = biwi.predict_files(hp, blob, off, 640, 480, K, ctx=ctx)
        ref = hp.predict_batch(frames, K, ctx=ctx)
    finally:
        ctx.set_chunk_frames(0)
    assert np.array_equal(got["mid_point"], ref["mid_point"]) and np.array_equal(got["rotation"], ref["rotation"])
    for i in (0, 5, 10):
        tr = of.predict(oracle.biwi_read_depth(files[i]), synth.KINECT_K, mode=oracle.MODE_SAT, keep=False)
        assert np.array_equal(got["mid_point"][i], tr.mid_point) and np.array_equal(got["rotation"][i], tr.rotation)
    # a damaged file in the batch is reported, not silently predicted on
    blob2, off2 = biwi.pack_files(files[:3] + [files[3][:100]] + files[4:6])
    with pytest.raises(DhError):
        biwi.predict_files(hp, blob2, off2, 640, 480, K, ctx=ctx)
    assert ctx.counters()["launches"] >= 0


def test_decode_mutated_streams_against_the_oracle(ctx):
    """Seeded mutation fuzzing of the run-length reader: bytes flipped, header fields replaced by
    extreme values, files cut short or extended.  For every stream the GPU reader either fails
    where the reference reader fails (biwi.rs:81-103: UnexpectedEof, index out of bounds) or
    returns the same pixels.  Both sides see the same bytes (the 16-byte padding of pack_files
    is part of the file the ABI is given)."""
    rng = np.random.default_rng(20240)
    lib = capi.load()
    n_fail = n_ok = 0
    for case in range(400):
        h, w = int(rng.integers(1, 49)), int(rng.integers(1, 65))
        if case % 50 == 0:
            h, w = 480, 640
        fill = float(rng.choice([0.02, 0.3, 0.5, 0.9]))
        frame = np.where(rng.random((h, w)) < fill, rng.integers(1, 65536, (h, w)), 0).astype(np.uint16)
        data = bytearray(biwi.encode_depth(frame))
        kind = int(rng.integers(0, 6))
        body = len(data) - 8
        if kind == 0 and body > 0:      # flip one byte behind the size header
            data[8 + int(rng.integers(0, body))] ^= int(rng.integers(1, 256))
        elif kind == 1 and body >= 4:   # an extreme value in some aligned word (a header field or two pixels)
            pos = 8 + 2 * int(rng.integers(0, (body - 4) // 2 + 1))
            data[pos:pos + 4] = struct.pack("<I", int(rng.choice([0, 1, w * h, w * h + 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0xFFFFFFFE])))
        elif kind == 2:                 # cut short
            data = data[: 8 + int(rng.integers(0, body + 1))]
        elif kind == 3:                 # trailing garbage
            data += bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))
        elif kind == 4 and body >= 8:   # delete an aligned piece
            pos = 8 + 2 * int(rng.integers(0, body // 2))
            del data[pos:pos + 2 * int(rng.integers(1, 5))]
        # kind 5: unchanged
        blob, off = biwi.pack_files([bytes(data)])
        seen = bytes(blob[: int(off[1])])
        try:
            want = oracle.biwi_read_depth(seen)
        except oracle.BiwiError:
            want = None
        out = np.full((1, h, w), 0xABCD, np.uint16)
        rc = lib.dh_biwi_decode_depth(ctx._h, capi.ptr(blob), capi.ptr(off), 1, w, h, capi.ptr(out), capi.DH_DEPTH_HOST)
        if want is None:
            assert rc == capi.DH_E_ARG, "case %d kind %d: the reference reader fails, the GPU reader returned %d" % (case, kind, rc)
            n_fail += 1
        else:
            assert rc == capi.DH_OK, "case %d kind %d: %s" % (case, kind, lib.dh_last_error())
            assert np.array_equal(out[0], want), "case %d kind %d" % (case, kind)
            n_ok += 1
    assert n_fail > 50 and n_ok > 50
    good = biwi.encode_depth(synth.make_frames(1, seed=4)[0])
    assert np.array_equal(biwi.read_depth(good, ctx=ctx), oracle.biwi_read_depth(good))


# ---- the compressed host->device path of dh_predict_batch (worker threads rewrite host frames as
#      run-length files, the GPU expands them): results must equal the raw copy and the device-resident pass
@pytest.mark.parametrize("env", [{"DH_HOST_ENCODE": "1", "DH_HOST_HYBRID": "0"}, {"DH_HOST_ENCODE": "1", "DH_LANES": "1"},
                                 {"DH_HOST_ENCODE": "1", "DH_ENCODE_THREADS": "1"}, {"DH_HOST_HYBRID": "0"}, {}])
def test_predict_batch_host_frames_through_the_run_length_rewrite(monkeypatch, env):
    import torch
    arr = synth.make_forest(seed=4, n_trees=3, max_depth=7)
    frames = synth.make_frames(44, seed=77)
    rng = np.random.default_rng(5)
    frames[3] = 0                                                       # all background
    frames[8] = rng.integers(1, 4000, frames[8].shape).astype(np.uint16)  # no background: the file is LARGER than the frame
    frames[17, ::2, ::3] = 0                                            # salt and pepper
    frames[40:] = rng.integers(0, 3, frames[40:].shape).astype(np.uint16) * 900   # a dense last chunk
    hp = HoughPrediction.from_arrays(arr, stepwidth=8)

    def run(e):
        for k in ("DH_HOST_ENCODE", "DH_LANES", "DH_ENCODE_THREADS", "DH_HOST_HYBRID"):
            monkeypatch.delenv(k, raising=False)
        for k, v in e.items():
            monkeypatch.setenv(k, v)
        c = Context(0)
        c.set_chunk_frames(10)
        try:
            out = hp.predict_batch(frames, K, ctx=c).copy()
            info = c.transfer_info()
            out2 = hp.predict_batch(frames[:25], K, ctx=c).copy()       # same context again, other batch size
        finally:
            c.close()
        return out, out2, info
    ref, ref2, info0 = run({"DH_HOST_ENCODE": "0"})
    assert info0["encoded_chunks"] == 0 and info0["h2d_bytes"] == frames.nbytes
    got, got2, info = run(env)
    for a, b in ((ref, got), (ref2, got2)):
        assert np.array_equal(a["mid_point"], b["mid_point"]) and np.array_equal(a["rotation"], b["rotation"])
    if env.get("DH_HOST_HYBRID") == "0" and env.get("DH_HOST_ENCODE") == "1":
        assert info["encoded_chunks"] == 5                              # every chunk waits for its rewrite
    elif env.get("DH_HOST_HYBRID") == "0":
        assert info["encoded_chunks"] == 4                              # the dense chunk goes raw
    else:
        assert 1 <= info["encoded_chunks"] <= 5                         # idle copy engine: chunks from the back go raw
    assert info["h2d_bytes"] < frames.nbytes
    # and the device-resident pass
    c = Context(0)
    try:
        dev = torch.from_numpy(frames.view(np.int16)).cuda()
        out = hp.predict_batch(None, K, ctx=c, device_ptr=dev.data_ptr(), n=len(frames), w=640, h=480)
    finally:
        c.close()
    assert np.array_equal(out["mid_point"], ref["mid_point"]) and np.array_equal(out["rotation"], ref["rotation"])

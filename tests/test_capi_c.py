"""The C ABI from C: include/depthhead_cuda.h must be a valid C99 header (it is what a Rust
`extern "C"` block, cgo or any other FFI binds), and a plain C program linked against
libdepthhead_cuda.so must be able to load a model, change its public scalars and — on a GPU — run
dh_predict / dh_predict_batch with the same results as the Python mirror."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from depthhead_b200 import Context, HoughPrediction, IntrinsicMatrix, capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "capi_smoke.c")


def _build(tmp_path):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    capi.load()  # builds the library if needed
    exe = str(tmp_path / "capi_smoke")
    libdir = os.path.dirname(capi.lib_path())
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", libdir, "-l:libdepthhead_cuda.so", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _model(tmp_path):
    arr = synth.make_forest(seed=3, n_trees=4, max_depth=6)
    js = synth.forest_to_json(arr, stepwidth=10)
    path = str(tmp_path / "model.json")
    with open(path, "w") as f:
        f.write(js)
    return arr, js, path


def test_header_is_c99_and_host_entry_points_work_from_c(tmp_path):
    exe = _build(tmp_path)
    _, _, model = _model(tmp_path)
    r = subprocess.run([exe, model], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    lines = r.stdout.strip().splitlines()
    assert lines[0].startswith("abi 1 build ") and lines[-1] == "ok"
    assert "trees 4 nodes 252 leaves 256" in lines[1] and "stepwidth 10 iterations 20 sigma 8" in lines[1]


def test_header_compiles_as_cxx_too(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    src = tmp_path / "inc.cpp"
    src.write_text('#include "depthhead_cuda.h"\nint main() { return sizeof(dh_result) == 56 ? 0 : 1; }\n')
    r = subprocess.run([gxx, "-std=c++11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_predict_from_c_equals_the_python_mirror(tmp_path):
    exe = _build(tmp_path)
    arr, js, model = _model(tmp_path)
    frame = synth.make_frames(1, seed=11)[0]
    fpath = str(tmp_path / "frame.u16")
    frame.tofile(fpath)
    r = subprocess.run([exe, model, fpath, "640", "480"], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    poses = [[float(x) for x in ln.split()[1:]] for ln in r.stdout.splitlines() if ln.startswith("pose ")]
    assert len(poses) == 5
    hp = HoughPrediction.from_json(js)
    K = IntrinsicMatrix.default_kinect_intrinsic()
    c = Context(0)
    try:
        a = hp.predict_parameter_parallel(frame, K, ctx=c)
        b = hp.predict_parameter_parallel(frame, K, [10.0, -20.0, 900.0], None, ctx=c)
    finally:
        c.close()
    for got, want in ((poses[0], a), (poses[1], b), (poses[2], a), (poses[3], a), (poses[4], a)):
        assert np.array_equal(np.float32(got[:3]), want.mid_point) and np.array_equal(np.float64(got[3:]), want.rotation)

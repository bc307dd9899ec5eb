import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The library keeps the patch gate inside the seed-grid kernel for passes of fewer than 64 frames (they are
# bound by the chain of launches) and runs it as its own kernel over probability codes from there on.  The
# tests work on a handful of frames: they take the large-batch path unless a test says otherwise
# (test_kernel_variants restores the library's threshold, test_large_batch_crosses_both_gate_paths uses both).
os.environ.setdefault("DH_GATE_SPLIT_MIN", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_case():
    """A small seeded forest + frames shared by several tests (depth-6 trees, 4 trees)."""
    from depthhead_b200 import synth
    arr = synth.make_forest(seed=3, n_trees=4, max_depth=6)
    js = synth.forest_to_json(arr, stepwidth=10)
    frames = synth.make_frames(3, seed=11)
    return arr, js, frames

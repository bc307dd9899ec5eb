import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


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

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


@pytest.fixture(scope="session")
def ref_sliding_window():
    """The reference's compiled SlidingWindow module (oracle/_ref), or skip."""
    from oracle import build_ref
    build_ref.build()
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref/SlidingWindow.*.so not built (reference tree absent)")
    return mod

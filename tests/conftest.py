import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_addoption(parser):
    parser.addoption("--emulate-gpu", action="store_true", default=False,
                     help="run the `-m gpu` tests on CPU tensors over the library's sources compiled for the host (tests/_fake_cuda.py)")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if config.getoption("--emulate-gpu"):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _fake_cuda
        _fake_cuda.enable()


def pytest_collection_modifyitems(config, items):
    if not config.getoption("--emulate-gpu"):
        return
    import _fake_cuda
    skip = pytest.mark.skip(reason="sized for the real machine: not run under --emulate-gpu")
    for item in items:
        if any(s in item.nodeid for s in _fake_cuda.TOO_LARGE):
            item.add_marker(skip)
        if any(s in item.nodeid for s in _fake_cuda.NOT_APPLICABLE):
            item.add_marker(pytest.mark.skip(reason="asserts that CPU tensors are rejected: not applicable under --emulate-gpu"))


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        return cache[name]

    return load

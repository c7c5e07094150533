import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ffr_lib():
    """Built + loaded C-ABI library (nvcc cross-compiles here without a GPU)."""
    from face_detection_and_recognition_b200 import _build, _lib
    if not os.path.exists(_build.LIB_PATH):
        _build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


class _FfrEnv:
    """FFR_* experiment knobs are read ONCE per process by libffr_b200.so; tests that flip one between calls go through
    this helper, which also asks the library to re-read its environment (non-ABI hook ffr_debug_reload_env)."""

    def __init__(self, lib):
        self._lib, self._old = lib, {}

    def setenv(self, key, value):
        self._old.setdefault(key, os.environ.get(key))
        os.environ[key] = str(value)
        self._lib.ffr_debug_reload_env()

    def delenv(self, key):
        self._old.setdefault(key, os.environ.get(key))
        os.environ.pop(key, None)
        self._lib.ffr_debug_reload_env()

    def restore(self):
        for k, v in self._old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        self._old.clear()
        self._lib.ffr_debug_reload_env()


@pytest.fixture
def ffr_env(ffr_lib):
    e = _FfrEnv(ffr_lib)
    yield e
    e.restore()

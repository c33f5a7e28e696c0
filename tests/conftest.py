import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "emul: runs the CUDA sources under the CPU thread emulator "
                                       "(developer aid, opt-in with SGX_EMUL_TESTS=1)")


_CACHE = {}


@pytest.fixture(scope="session")
def recordings():
    """name -> (spec, int8 data); generated lazily once per session."""
    from tests.cases import CASES, build_recording

    class Lazy(dict):
        def __missing__(self, k):
            self[k] = build_recording(CASES[k])
            return self[k]
    return Lazy()

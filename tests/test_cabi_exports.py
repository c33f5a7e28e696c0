"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol declared in
include/softgnss_b200.h, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from softgnss_python_b200 import build, _native
    build.build_native()
    return _native.Lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "softgnss_b200.h")).read()
    declared = set(re.findall(r"\b(sgx_[a-z_]+)\s*\(", hdr))
    assert declared >= {"sgx_acquire", "sgx_track", "sgx_synth_generate", "sgx_device_count"}
    for name in sorted(declared):
        assert hasattr(lib.dll, name), name
    assert lib.dll.sgx_abi_version() == 1


def test_struct_layouts_match_header(lib):
    from softgnss_python_b200 import _native
    from softgnss_python_b200.settings import SgxSettings
    assert ctypes.sizeof(_native.SgxChannel) == 24
    assert ctypes.sizeof(SgxSettings) == 13 * 8 + 8 + 10 * 4
    assert ctypes.sizeof(_native.SgxSynthSpec) == 8 + 4 * 4 + 3 * 12 * 4 + 4 * 12 * 8
    assert ctypes.sizeof(_native.SgxNavSettings) == 5 * 8 + 2 * 4 and len(_native.EPH_FIELDS) == 21   # sgx_eph: 21 doubles


def test_no_cpu_fallback(lib):
    """Without a GPU the product path must fail loudly, never compute on the host."""
    if lib.dll.sgx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from softgnss_python_b200 import _native
    from softgnss_python_b200.acquisition import acquisition
    from softgnss_python_b200.settings import Settings
    from softgnss_python_b200.tracking import tracking
    with pytest.raises(_native.NativeError):
        acquisition(np.zeros(11 * 38192, dtype=np.int8), Settings())
    ch = np.rec.fromarrays([[1], [9.548e6], [0.0], ['T']], names="PRN,acquiredFreq,codePhase,status")
    with pytest.raises(_native.NativeError):
        tracking(np.zeros(3 * 38192, dtype=np.int8), ch, Settings(numberOfChannels=1, msToProcess=1.0))
    from softgnss_python_b200 import postnav
    with pytest.raises(_native.NativeError):
        postnav.find_preambles_batch(np.ones((1, 8000)))
    s = Settings(numberOfChannels=4, msToProcess=2000.0)
    with pytest.raises(_native.NativeError):
        postnav.nav_solve_batch(np.zeros((1, 4, 2000)), np.zeros((1, 4)), np.ones((1, 4)), np.zeros((1, 4, 21)), [0.0], s)

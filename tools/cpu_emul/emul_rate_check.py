"""Developer check: acquisition at another sampling rate under the CPU fiber emulator vs the oracle
(scenario of tests/test_gpu_configs.py::test_sampling_rate_sweep_acquire_and_track)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native                      # noqa: E402
_native.LIB_PATH = os.path.join(ROOT, "tools", "cpu_emul", "libsoftgnss_emul.so")
_native._LIB = _native.Lib(_native.LIB_PATH)
from softgnss_python_b200.acquisition import acquire_batch    # noqa: E402
from oracle import gnss_oracle as orc                          # noqa: E402
from tests.test_gpu_configs import _scenario                   # noqa: E402

fs = float(sys.argv[1]) if len(sys.argv) > 1 else 16.3676e6
f_if = float(sys.argv[2]) if len(sys.argv) > 2 else 4.1304e6
s, data, nlong, nprn = _scenario(fs, f_if, ms=12)
ref = orc.acquire(data[:nlong], s, clamp_window=True)
got = acquire_batch(data[:nlong].reshape(1, -1), s)
print("detected", got["carrFreq"][0] > 0, ref["carrFreq"][:nprn] > 0)
print("codePhase", got["codePhase"][0], ref["codePhase"][:nprn])
print("carrFreq diff", np.abs(got["carrFreq"][0] - ref["carrFreq"][:nprn]).max())
print("peakMetric rel", np.abs(got["peakMetric"][0] / ref["peakMetric"][:nprn] - 1).max())

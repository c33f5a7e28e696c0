"""Developer check: acquisition pipeline under the CPU fiber emulator vs the reference's golden."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native                      # noqa: E402
_native.LIB_PATH = os.path.join(ROOT, "tools", "cpu_emul", "libsoftgnss_emul.so")
_native._LIB = _native.Lib(_native.LIB_PATH)
from softgnss_python_b200.acquisition import acquire_batch    # noqa: E402
from tests.cases import CASES, N, build_recording, case_settings   # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "acq_c1"
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 3
case = CASES[name]
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
spec, data = build_recording(case)
s = case_settings(case)
t = time.time()
skip = s.skipNumberOfBytes
r = acquire_batch(data[skip:skip + 11 * N].reshape(1, -1), s, prn_first=first, prn_count=count, diagnostics=True)
print("%.1fs" % (time.time() - t))
sl = slice(first, first + count)
print("peakMetric got", r["peakMetric"][0])
print("peakMetric ref", g["peakMetric"][sl], "rel", np.abs(r["peakMetric"][0] / g["peakMetric"][sl] - 1).max())
print("codePhase got", r["codePhase"][0], "ref", g["codePhase"][sl])
print("carrFreq got", r["carrFreq"][0], "ref", g["carrFreq"][sl])
print("bins", r["frqBin"][0], "fine", r["finePeakIndex"][0])

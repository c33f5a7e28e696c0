import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softgnss_python_b200.acquisition import acquire_batch
from tests.cases import CASES, N, build_recording, case_settings
case = CASES["acq_c1"]
g = np.load(os.path.join(ROOT, "tests", "golden", "acq_c1.npz"))
spec, data = build_recording(case)
s = case_settings(case)
nprn = int(sys.argv[1]) if len(sys.argv) > 1 else 32
s.acqSatelliteList = range(1, nprn + 1)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    r = acquire_batch(data[:11 * N].reshape(1, -1), s, diagnostics=True)
    det = r["carrFreq"][0] > 0
    print("run", it, "fine idx", r["finePeakIndex"][0][det].tolist())
    print("   ref     ", np.round(g["carrFreq"][:nprn][g["carrFreq"][:nprn] > 0] * (8 * 2 ** 19) / 38192000.0).astype(int).tolist())
    print("   metric rel", np.abs(r["peakMetric"][0] / g["peakMetric"][:nprn] - 1).max())

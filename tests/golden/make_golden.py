"""Generate the golden vectors under tests/golden/ by running the REFERENCE itself.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference through the mechanical Python-3 shim (oracle/make_ref_shim.py,
SURVEY.md appendix B), feeds it seeded synthetic recordings from
``softgnss_python_b200.synth`` and stores the reference's outputs.  The inputs are not
stored (they are regenerated from the seed; their SHA-1 is stored and checked).
"""
import hashlib
import io
import os
import sys
import contextlib
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import make_ref_shim                      # noqa: E402
from softgnss_python_b200 import synth                # noqa: E402
from tests.cases import CASES, build_recording       # noqa: E402


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = make_ref_shim.import_ref()
    init, acq_mod, trk_mod = ref["initialize"], ref["acquisition"], ref["tracking"]

    # ---- known answers of the signal-definition helpers -------------------------------
    s = init.Settings()
    codes = np.array([s.generateCAcode(p) for p in range(32)])
    table = s.makeCaTable()
    np.savez_compressed(
        os.path.join(HERE, "helpers.npz"),
        codes_sha1=sha1(codes.astype(np.int8)), table_sha1=sha1(table.astype(np.int8)),
        prn1_first10=codes[0, :10], chip_sums=codes.sum(1),
        dll_coef=np.array(s.calcLoopCoef(2.0, 0.7, 1.0)),
        pll_coef=np.array(s.calcLoopCoef(25.0, 0.7, 0.25)),
        samples_per_code=s.samplesPerCode)

    for name, case in CASES.items():
        spec, data = build_recording(case)
        s = init.Settings()
        for k, v in case.get("settings", {}).items():
            setattr(s, k, v)
        s.msToProcess = float(case["ms"])
        n = s.samplesPerCode
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            a = acq_mod.AcquisitionResult(s)
            a.acquire(data[s.skipNumberOfBytes:s.skipNumberOfBytes + 11 * n])
            a.preRun()
        out = dict(input_sha1=sha1(data), carrFreq=a.carrFreq, codePhase=a.codePhase,
                   peakMetric=a.peakMetric, ch_PRN=a.channels.PRN,
                   ch_acquiredFreq=a.channels.acquiredFreq, ch_codePhase=a.channels.codePhase,
                   ch_status=np.array([str(x) for x in a.channels.status]))
        if case["ms"] > 0:
            t = trk_mod.TrackingResult(a)
            with tempfile.NamedTemporaryFile(suffix=".bin") as tf:
                data.tofile(tf.name)
                with contextlib.redirect_stdout(sink), open(tf.name, "rb") as fid:
                    t.track(fid)
            r = t.results
            out["trk_PRN"] = np.array(r.PRN)
            for f in ("absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P",
                      "Q_L", "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt"):
                out["trk_" + f] = np.stack([np.asarray(x, dtype=np.float64) for x in r[f]])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "PRNs detected:", (np.nonzero(a.carrFreq)[0] + 1).tolist(),
              "truth:", [int(p) for p in spec.prn])


if __name__ == "__main__":
    main()

"""SURVEY.md section 7 step 7 as a test: the reference's own caller, `initialize.Settings.postProcessing`
(initialize.py:454-527), run UNCHANGED (mechanical Python-3 shim only) on top of the drop-in modules: `import acquisition`
and `import tracking` resolve to `dropin/` (the B200 path), `postNavigation` stays the reference's own.  The file is the
37 100 ms LNAV recording of config 2; the navigation solutions the reference computes from the B200 tracking results must
equal those it computed from its own tracking (tests/golden/c2_full.npz, made by the all-reference run).

Needs a GPU and the shimmed reference: `/root/reference` (build container) or the generated scratch copy `oracle/_ref/`
(git-ignored; it travels with a gpurun snapshot).  Skipped when neither is present."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_postProcessing_on_the_dropins(tmp_path):
    have_ref = os.path.isdir("/root/reference") or os.path.exists(os.path.join(ROOT, "oracle", "_ref", "initialize.py"))
    if not have_ref:
        pytest.skip("the shimmed reference is not available here (no /root/reference, no oracle/_ref)")
    out = str(tmp_path / "nav.npz")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_caller_worker.py"), out, str(tmp_path)],
                         capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert res.returncode == 0, (res.stdout[-3000:], res.stderr[-3000:])
    got = np.load(out, allow_pickle=False)
    g = np.load(os.path.join(ROOT, "tests", "golden", "c2_full.npz"), allow_pickle=False)
    assert np.array_equal(got["ch_PRN"], g["ch_PRN"]) and np.array_equal(got["codePhase"], g["codePhase"])
    assert np.abs(got["carrFreq"] - g["carrFreq"]).max() <= 1.0
    n = int(np.sum(~np.isnan(g["sol_X"])))
    assert n > 50
    # the B200 acquisition may report a carrier a fine-FFT bin away from the reference's (<= 1 Hz contract); tracking
    # then converges to the same lock, and the downstream consumer must produce the same fixes
    for f, tol in (("X", 1e-3), ("Y", 1e-3), ("Z", 1e-3), ("dt", 1e-3), ("height", 1e-3), ("E", 1e-3), ("N", 1e-3), ("U", 1e-3),
                   ("latitude", 1e-8), ("longitude", 1e-8)):
        a, b = got[f][:n], g["sol_" + f][:n]
        assert np.array_equal(np.isnan(a), np.isnan(b)), f
        assert np.nanmax(np.abs(a - b)) <= tol, "%s differs by %.3g" % (f, np.nanmax(np.abs(a - b)))
    assert np.nanmax(np.abs(got["rawP"][:, :n] - g["solch_rawP"][:, :n])) <= 1e-3
    rec = dict(fixes=n, max_abs_diff_m=float(max(np.nanmax(np.abs(got[f][:n] - g["sol_" + f][:n])) for f in "XYZ")),
               absoluteSample_identical=bool(np.array_equal(
                   got["absoluteSample"],
                   (g["ch_codePhase"][:8, None] + np.cumsum(g["abs_delta"].astype(np.float64) + 38192, axis=1)))))
    print(rec)
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        import json
        with open(os.path.join(ROOT, "gpurun_out", "reference_caller.json"), "w") as fh:
            json.dump(rec, fh)

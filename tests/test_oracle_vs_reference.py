"""Live diff of the oracle restatement against the reference itself (through the Python-3 shim) on
inputs the golden set does not contain.  Runs only where /root/reference exists (the build
container); skipped on the GPU box."""
import contextlib
import io
import os
import tempfile

import numpy as np
import pytest

REF = os.environ.get("SGX_REFERENCE_DIR", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def test_fresh_seed_acquire_prerun_track_bit_identical():
    from oracle import gnss_oracle as orc
    from oracle import make_ref_shim
    from softgnss_python_b200 import synth
    ref = make_ref_shim.import_ref()
    n = 38192
    sats = synth.default_constellation(seed=4242, n_sats=3, cn0=50.0)
    spec = synth.RecordingSpec(sats, seed=4242)
    ms = 40
    data = synth.generate_cpu(spec, (ms + 12) * n)
    s = ref["initialize"].Settings()
    s.msToProcess = float(ms)
    s.numberOfChannels = 3
    s.acqSatelliteList = range(1, 33)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        a = ref["acquisition"].AcquisitionResult(s)
        a.acquire(data[:11 * n])
        a.preRun()
        t = ref["tracking"].TrackingResult(a)
        with tempfile.NamedTemporaryFile(suffix=".bin") as tf:
            data.tofile(tf.name)
            with open(tf.name, "rb") as fid:
                t.track(fid)
    mine = orc.acquire(data[:11 * n], s)
    for f in ("carrFreq", "codePhase", "peakMetric"):
        assert np.array_equal(mine[f], getattr(a, f)), f
    ch = orc.pre_run(mine, s)
    assert np.array_equal(ch["PRN"], a.channels.PRN)
    recs = orc.track(data, ch, s)
    for i, (prn, status, series) in enumerate(recs):
        assert prn == t.results[i].PRN
        for f in orc.TRACK_FIELDS:
            assert np.array_equal(series[f], np.asarray(t.results[i][f], dtype=np.float64)), f

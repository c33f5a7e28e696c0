"""GPU parity on BASELINE.json config 3 AS SPECIFIED: weak-signal acquisition with 10 ms coherent integration and with
ten 1 ms blocks, 100 Hz Doppler step (141 bins), all 32 PRNs, satellites at 30-43 dB-Hz -- including PRNs whose
peakMetric lies within 10 % of acqThreshold, where a float32-vs-float64 difference would flip the decision.

The reference hard-codes two 1 ms blocks and 500 Hz (acquisition.py:55-57, :68, :101, :129-133); the extension is pinned
by the oracle's parametrised restatement, which needs minutes per recording at these sizes, hence the committed fixture
tests/golden/acq_c3.npz (made by tests/golden/make_golden_c3.py).  Tolerances: detected set, codePhase exact;
carrFreq <= 1 Hz; peakMetric 1e-5 relative (float32 transforms against the oracle's float64)."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "acq_c3.npz")
N = 38192


def _recording(seed):
    from softgnss_python_b200 import synth
    from tests.golden.make_golden_c3 import recording_spec
    return synth.generate_cpu(recording_spec(seed), 11 * N)


def _check(got, g, key, row=0):
    ref_c, ref_p, ref_m = g[key + "carrFreq"], g[key + "codePhase"], g[key + "peakMetric"]
    assert np.array_equal(got["carrFreq"][row] > 0, ref_c > 0), key + " detected PRN set"
    assert np.array_equal(got["codePhase"][row], ref_p), key + " codePhase"
    assert np.abs(got["carrFreq"][row] - ref_c).max() <= 1.0, key + " carrFreq"
    assert np.abs(got["peakMetric"][row] / ref_m - 1).max() <= 1e-5, key + " peakMetric"


@pytest.mark.parametrize("mode", ["coh10", "blk10"])
@pytest.mark.parametrize("seed", [1000, 1001])
def test_config3_at_spec(mode, seed):
    from softgnss_python_b200.acquisition import acquire_batch
    from tests.golden.make_golden_c3 import settings_for
    g = np.load(GOLD, allow_pickle=False)
    key = "%s_%d_" % (mode, seed)
    data = _recording(seed)
    assert hashlib.sha1(data.tobytes()).hexdigest() == str(g[key + "sha1"])
    s = settings_for(mode)
    got = acquire_batch(data.reshape(1, -1), s)
    assert got["carrFreq"].shape == (1, 32)
    _check(got, g, key)
    # the fixture is only meaningful if it contains decisions next to the threshold
    m = g[key + "peakMetric"]
    assert (m > s.acqThreshold).any() and (m < s.acqThreshold).any()


def test_config3_has_decisions_within_ten_percent_of_the_threshold():
    g = np.load(GOLD, allow_pickle=False)
    near = [float(v) for k in g.files if k.endswith("peakMetric") for v in g[k] if 2.25 <= v <= 2.75]
    assert len(near) >= 3 and min(near) < 2.5 < max(near)


@pytest.mark.parametrize("mode", ["coh10", "blk10"])
def test_config3_batch_and_prn_shards(mode):
    """A batch of recordings walked in slices (acquire_batch splits large batches) and a PRN-sharded search return
    the bytes of the single-recording search."""
    from softgnss_python_b200 import acquisition
    from tests.golden.make_golden_c3 import settings_for
    g = np.load(GOLD, allow_pickle=False)
    s = settings_for(mode)
    d0, d1 = _recording(1000), _recording(1001)
    batch = np.stack([d0, d1, d0])
    old = acquisition.ACQ_SPECTRA_BYTES
    try:
        acquisition.ACQ_SPECTRA_BYTES = 1 << 29           # forces one recording per sgx_acquire call
        got = acquisition.acquire_batch(batch, s)
    finally:
        acquisition.ACQ_SPECTRA_BYTES = old
    _check(got, g, "%s_1000_" % mode, 0)
    _check(got, g, "%s_1001_" % mode, 1)
    for f in ("carrFreq", "codePhase", "peakMetric"):
        assert np.array_equal(got[f][0], got[f][2])
    whole = acquisition.acquire_batch(batch[:2], s)
    for f in ("carrFreq", "codePhase", "peakMetric"):
        assert np.array_equal(whole[f], got[f][:2]), f + " batch invariance"
    parts = [acquisition.acquire_batch(batch[:2], s, prn_first=lo, prn_count=hi - lo) for lo, hi in ((0, 11), (11, 22), (22, 32))]
    for f in ("carrFreq", "codePhase", "peakMetric"):
        assert np.array_equal(np.concatenate([p[f] for p in parts], axis=1), whole[f]), f + " PRN shards"

"""CPU: the oracle restatement against the config-2 golden made by the reference's own caller
(tests/golden/c2_full.npz, tests/golden/make_golden_c2.py): acquisition of the first 11 ms and the first 75 ms of
tracking of all eight channels must reproduce the reference's stored values BIT FOR BIT (the recording is regenerated
from the seed by the CPU twin of the device generator), and the config-3 fixture must be self-consistent."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")
N = 38192


def test_oracle_reproduces_the_head_of_the_config2_golden():
    from oracle import gnss_oracle as orc
    from softgnss_python_b200 import navsynth, synth
    from softgnss_python_b200.settings import Settings
    g = np.load(os.path.join(GOLD, "c2_full.npz"), allow_pickle=False)
    spec, _ = navsynth.build_scenario(seed=int(g["seed"]))
    ms = 75
    data = synth.generate_cpu(spec, (ms + 3) * N)
    s = Settings(msToProcess=float(ms), numberOfChannels=8)
    acq = orc.acquire(data[:11 * N], s)
    for f in ("carrFreq", "codePhase", "peakMetric"):
        assert np.array_equal(acq[f], g[f]), f
    ch = orc.pre_run(acq, s)
    assert np.array_equal(ch["PRN"], g["ch_PRN"]) and np.array_equal(ch["acquiredFreq"], g["ch_acquiredFreq"])
    recs = orc.track(data, ch, s)
    sub = int(g["sub"])
    start = g["ch_codePhase"][:8]
    ref_abs = start[:, None] + np.cumsum(g["abs_delta"][:, :ms].astype(np.float64) + N, axis=1)
    sign = np.unpackbits(g["ip_sign"], axis=1)[:, :ms].astype(bool)
    for c, r in enumerate(recs):
        series = r[2]
        assert np.array_equal(series["absoluteSample"], ref_abs[c]), "absoluteSample, channel %d" % c
        assert np.array_equal(series["I_P"] > 0, sign[c]), "sign(I_P), channel %d" % c
        for i, f in enumerate(orc.TRACK_FIELDS):
            assert np.array_equal(series[f][::sub], g["sub_series"][c, i, :(ms + sub - 1) // sub]), "%s, channel %d" % (f, c)


def test_config3_fixture_is_complete():
    g = np.load(os.path.join(GOLD, "acq_c3.npz"), allow_pickle=False)
    for mode in ("coh10", "blk10"):
        for seed in (1000, 1001):
            k = "%s_%d_" % (mode, seed)
            assert g[k + "peakMetric"].shape == (32,) and len(str(g[k + "sha1"])) == 40
            det = g[k + "carrFreq"] > 0
            assert np.array_equal(det, g[k + "peakMetric"] > 2.5)
            assert np.all(g[k + "codePhase"][~det] == 0)

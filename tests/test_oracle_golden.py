"""The oracle restatement (oracle/gnss_oracle.py) and the host-side helpers against the golden
vectors produced by the reference itself (tests/golden/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import gnss_oracle as orc
from softgnss_python_b200 import settings as st
from tests.cases import CASES, N, case_settings

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def test_helpers_known_answers():
    g = gold("helpers")
    s = st.Settings()
    assert s.samplesPerCode == int(g["samples_per_code"]) == 38192
    codes = np.array([s.generateCAcode(p) for p in range(32)])
    assert codes.dtype == np.float64 and set(np.unique(codes)) == {-1.0, 1.0}
    assert sha1(codes.astype(np.int8)) == str(g["codes_sha1"]) == "b5d7cd36e0d35ca34e5f114ead37e57ebf477059"
    assert np.array_equal(codes[0, :10], g["prn1_first10"])        # ICD-200 octal 1440 for PRN 1
    assert np.array_equal(codes.sum(1), g["chip_sums"])
    assert sha1(s.makeCaTable().astype(np.int8)) == str(g["table_sha1"])
    assert np.array_equal(np.array(s.calcLoopCoef(2.0, 0.7, 1.0)), g["dll_coef"])
    assert np.array_equal(np.array(s.calcLoopCoef(25.0, 0.7, 0.25)), g["pll_coef"])
    # the oracle's own (+-1 register) generator agrees with the bit-level one
    assert np.array_equal(np.array([orc.ca_code(p) for p in (0, 6, 18, 31)]), codes[[0, 6, 18, 31]])
    assert np.array_equal(orc.ca_table(s), s.makeCaTable())


def test_prn_out_of_range_asserts():
    with pytest.raises(AssertionError):
        st.Settings().generateCAcode(32)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_outputs(name, recordings):
    g = gold(name)
    case = CASES[name]
    spec, data = recordings[name]
    assert sha1(data) == str(g["input_sha1"]), "synthetic generator drifted from the golden input"
    s = case_settings(case)
    skip = s.skipNumberOfBytes
    acq = orc.acquire(data[skip:skip + 11 * N], s)
    # same numpy calls in the same order: bit-identical, not merely close
    for f in ("carrFreq", "codePhase", "peakMetric"):
        assert np.array_equal(acq[f], g[f]), f
    ch = orc.pre_run(acq, s)
    assert np.array_equal(ch["PRN"], g["ch_PRN"])
    assert np.array_equal(ch["acquiredFreq"], g["ch_acquiredFreq"])
    assert np.array_equal(ch["codePhase"], g["ch_codePhase"])
    assert list(ch["status"]) == [str(x) for x in g["ch_status"]]
    if case["ms"] == 0:
        return
    recs = orc.track(data, ch, s)
    assert [r[0] for r in recs] == g["trk_PRN"].tolist()
    for f in orc.TRACK_FIELDS:
        got = np.stack([r[2][f] for r in recs])
        assert np.array_equal(got, g["trk_" + f]), f


def test_short_recording_returns_none(recordings):
    """tracking.py:159-163: a short read aborts the whole call."""
    case = CASES["trk_skip"]
    _, data = recordings["trk_skip"]
    s = case_settings(case)
    g = gold("trk_skip")
    ch = dict(PRN=g["ch_PRN"], acquiredFreq=g["ch_acquiredFreq"], codePhase=g["ch_codePhase"],
              status=[str(x) for x in g["ch_status"]])
    assert orc.track(data[:60 * N], ch, s) is None

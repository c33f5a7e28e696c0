"""Preamble search / bit summation: the oracle restatement against the output of the
reference's own findPreambles (tests/golden/make_golden_bitsync.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import gnss_oracle as orc
from tests.cases import BITSYNC_CHANNELS, BITSYNC_EARLY, build_bitsync_case, build_bitsync_channel

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bitsync.npz")


@pytest.fixture(scope="module")
def case():
    ips = build_bitsync_case()
    g = np.load(GOLD, allow_pickle=False)
    assert hashlib.sha1(np.ascontiguousarray(np.stack(ips)).tobytes()).hexdigest() == str(g["input_sha1"])
    return ips, g


def test_find_preambles_matches_reference(case):
    ips, g = case
    first, active = orc.find_preambles(ips)
    assert np.array_equal(first, g["first"])
    assert np.array_equal(active, g["active"])
    # what each constructed channel is meant to exercise
    by = {c["name"]: int(first[i]) for i, c in enumerate(BITSYNC_CHANNELS)}
    assert by["clean"] == 3655 and by["inverted"] == 977 and by["late"] == 41 and by["zeros"] == 5999
    assert by["random"] == 0 and by["weak"] == 0
    assert by["noisy"] == 2222 + 6000              # first preamble below the 153 threshold
    assert by["tlm_parity_broken"] == 1500 + 6000  # first preamble fails navPartyChk


def test_nav_bits_match_reference(case):
    ips, g = case
    bits = np.unpackbits(g["nav_bits"], axis=1)[:, :1501]
    n = 0
    for ch in range(len(ips)):
        if g["bits_valid"][ch]:
            assert np.array_equal(orc.nav_bits(ips[ch], int(g["first"][ch])), bits[ch]), ch
            n += 1
    assert n >= 4
    zi = [c["name"] for c in BITSYNC_CHANNELS].index("zeros")
    assert bits[zi][1 + 70] == 0                   # the bit whose 20 ms sum is exactly 0.0


def test_parity_check_known_answers():
    from softgnss_python_b200 import navsynth
    word = navsynth.lnav_word([1, 0] * 12, 0, 1)                       # a valid word after D29*=0, D30*=1
    nd = np.array([-1, 1] + [2 * b - 1 for b in word])
    assert orc.nav_party_chk(nd) == -1
    nd[5] *= -1
    assert orc.nav_party_chk(nd) == 0


def test_candidate_too_close_to_the_start():
    ip = build_bitsync_channel(BITSYNC_EARLY, 600)
    first, active = orc.find_preambles([ip])
    assert first[0] == 6020 and list(active) == [0]
    with pytest.raises(IndexError):
        orc.find_preambles([ip], skip_unreadable=False)


def test_pseudoranges_match_reference():
    from tests.cases import build_pseudo_case
    abs_sample, ms_index, act = build_pseudo_case()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pseudo.npz"), allow_pickle=False)
    assert hashlib.sha1(np.ascontiguousarray(abs_sample).tobytes()).hexdigest() == str(g["input_sha1"])
    for e in range(ms_index.shape[0]):
        got = orc.calculate_pseudoranges(abs_sample, ms_index[e], act[e].nonzero()[0], abs_sample.shape[0], 38192)
        assert np.array_equal(got, g["pseudoranges"][e], equal_nan=True), e
    assert np.isinf(g["pseudoranges"][1][2]) and np.isnan(g["pseudoranges"][3]).all()

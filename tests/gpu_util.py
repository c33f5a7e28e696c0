"""Helpers shared by the -m gpu parity tests."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")

# Tolerances.  The correlator sums are float32 products accumulated in float64 (observed 2e-7 of full
# scale against the reference's float64); the loops are contractive, so that error does not grow
# (SURVEY.md appendix E).  TIGHT applies as long as every sample is assigned to the same chip as in the
# reference.  The reference assigns chips with ceil() of a float64 code phase (tracking.py:166-188):
# when a chip boundary falls within ~1e-7 samples of a sampling instant (families of 341 boundaries
# sweep across the sample grid every few ms) a 1e-9-chip difference in the carried code phase moves
# ONE sample to the neighbouring chip.  That changes a correlator by at most 2*|sample| and kicks the
# loops by a bounded amount that decays with the loop bandwidth (SURVEY.md appendix A.2-4 anticipates
# exactly this).  From the first such reassignment on, LOOSE applies; decisions stay exact throughout.
TIGHT = dict(IQ_REL=1e-5, CARR_HZ=1e-3, CODE_HZ=1e-4, DISCR=1e-5)
LOOSE = dict(IQ_ABS=4 * 128.0, CARR_HZ=1.0, CODE_HZ=0.05, DLL=0.05, PLL=0.01)
TOL = TIGHT


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def channels_from_gold(g):
    return np.rec.fromarrays([g["ch_PRN"], g["ch_acquiredFreq"], g["ch_codePhase"],
                              [str(x) for x in g["ch_status"]]],
                             names="PRN,acquiredFreq,codePhase,status")


def compare_tracking(got, ref, label="", strict=False):
    """got/ref: dict field -> float64 [channels, ms].  Bit-exact decisions; TIGHT tolerances up to the
    first single-sample chip reassignment of a channel, LOOSE after it.  Returns the per-channel index of
    the first reassignment (== ms when there is none); strict=True forbids any."""
    iq = ("I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L")
    assert np.array_equal(np.sign(got["I_P"]), np.sign(ref["I_P"])), label + " nav-bit signs"
    assert np.abs(got["absoluteSample"] - ref["absoluteSample"]).max() <= 1, label + " absoluteSample"
    scale = max(np.abs(ref[f]).max() for f in iq)
    nch, ms = ref["I_P"].shape
    firsts = []
    for c in range(nch):
        dev = np.max([np.abs(got[f][c] - ref[f][c]) for f in iq], axis=0)
        bad = np.nonzero(dev > TIGHT["IQ_REL"] * scale)[0]
        first = int(bad[0]) if len(bad) else ms
        firsts.append(first)
        pre = slice(0, first)
        tag = "%s ch%d " % (label, c)
        assert np.array_equal(got["absoluteSample"][c, pre], ref["absoluteSample"][c, pre]), tag + "absoluteSample"
        for f, tol in (("carrFreq", TIGHT["CARR_HZ"]), ("pllDiscrFilt", TIGHT["CARR_HZ"]),
                       ("codeFreq", TIGHT["CODE_HZ"]), ("dllDiscrFilt", TIGHT["CODE_HZ"]),
                       ("dllDiscr", TIGHT["DISCR"]), ("pllDiscr", TIGHT["DISCR"])):
            if first > 0:
                assert np.abs(got[f][c, pre] - ref[f][c, pre]).max() <= tol, tag + f + " (tight)"
        if first < ms:
            assert not strict, tag + "chip reassignment at ms %d (dev %.3g)" % (first, dev[first])
            assert dev[first] <= LOOSE["IQ_ABS"], tag + "first deviation is larger than two samples"
            assert dev.max() <= LOOSE["IQ_ABS"] + TIGHT["IQ_REL"] * scale, tag + "I/Q (loose)"
            for f, tol in (("carrFreq", LOOSE["CARR_HZ"]), ("pllDiscrFilt", LOOSE["CARR_HZ"]),
                           ("codeFreq", LOOSE["CODE_HZ"]), ("dllDiscrFilt", LOOSE["CODE_HZ"]),
                           ("dllDiscr", LOOSE["DLL"]), ("pllDiscr", LOOSE["PLL"])):
                assert np.abs(got[f][c] - ref[f][c]).max() <= tol, tag + f + " (loose)"
    return firsts

"""Helpers shared by the -m gpu parity tests."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")

# Tolerances (float32 correlator sums perturb the loops by ~1e-6 relative;
# the loops are contractive, so the error does not accumulate).
TOL = dict(IQ_REL=1e-5, CARR_HZ=1e-3, CODE_HZ=1e-4, DISCR=1e-5)


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def channels_from_gold(g):
    return np.rec.fromarrays([g["ch_PRN"], g["ch_acquiredFreq"], g["ch_codePhase"],
                              [str(x) for x in g["ch_status"]]],
                             names="PRN,acquiredFreq,codePhase,status")


def compare_tracking(got, ref, label=""):
    """got/ref: dict field -> float64 [channels, ms].  Bit-exact where the domain is integer,
    stated tolerances elsewhere."""
    assert np.array_equal(got["absoluteSample"], ref["absoluteSample"]), label + " absoluteSample"
    assert np.array_equal(np.sign(got["I_P"]), np.sign(ref["I_P"])), label + " nav-bit signs"
    scale = max(np.abs(ref[f]).max() for f in ("I_P", "Q_P", "I_E", "Q_E", "I_L", "Q_L"))
    for f in ("I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L"):
        assert np.abs(got[f] - ref[f]).max() <= TOL["IQ_REL"] * scale, label + " " + f
    assert np.abs(got["carrFreq"] - ref["carrFreq"]).max() <= TOL["CARR_HZ"], label + " carrFreq"
    assert np.abs(got["pllDiscrFilt"] - ref["pllDiscrFilt"]).max() <= TOL["CARR_HZ"], label
    assert np.abs(got["codeFreq"] - ref["codeFreq"]).max() <= TOL["CODE_HZ"], label + " codeFreq"
    assert np.abs(got["dllDiscrFilt"] - ref["dllDiscrFilt"]).max() <= TOL["CODE_HZ"], label
    assert np.abs(got["dllDiscr"] - ref["dllDiscr"]).max() <= TOL["DISCR"], label + " dllDiscr"
    assert np.abs(got["pllDiscr"] - ref["pllDiscr"]).max() <= TOL["DISCR"], label + " pllDiscr"

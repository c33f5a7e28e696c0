"""LNAV encoder / geometry used for full-pipeline recordings.  CPU only.
The encoder is the inverse of what the reference decodes; where the reference tree is present its own
parity check and ephemeris decoder are run on the encoded bits."""
import os
import warnings

import numpy as np
import pytest

from softgnss_python_b200 import navsynth, synth

REF = os.environ.get("SGX_REFERENCE_DIR", "/root/reference")


def _parity_ok(prev2, word30):
    """IS-GPS-200 parity check of one received word given D29*, D30* (independent restatement)."""
    d29s, d30s = prev2
    d = [b ^ d30s for b in word30[:24]]
    return navsynth.lnav_word(d, d29s, d30s) == list(word30)


def test_words_pass_parity_and_carry_preamble():
    q = navsynth.quantize_ephemeris(dict(M_0=0.3, e=0.004, sqrtA=5153.7, omega_0=-1.0, i_0=0.96, omega=0.5, t_oe=388800.0))
    bits = navsynth.encode_stream(q, 4, 388794, 7).tolist()
    assert len(bits) == 2100
    for w in range(1, 70):
        assert _parity_ok(bits[30 * w - 2:30 * w], bits[30 * w:30 * w + 30]), w
    for sf in range(7):
        word = bits[300 * sf:300 * sf + 8]
        d30s = bits[300 * sf - 1] if sf else 0
        assert [b ^ d30s for b in word] == list(navsynth.PREAMBLE)


def test_quantisation_roundtrip_and_orbit_radius():
    e = dict(M_0=-2.0, e=0.0051, sqrtA=5154.5, omega_0=-0.58, i_0=0.969, omega=-1.9, omegaDot=-8e-9, deltan=4.5e-9,
             t_oe=388800.0)
    eq = navsynth.dequantize_ephemeris(navsynth.quantize_ephemeris(e))
    for k in ("M_0", "omega_0", "i_0", "omega"):
        assert abs(eq[k] - e[k]) < 2e-9
    assert abs(eq["e"] - e["e"]) < 2e-10 and abs(eq["sqrtA"] - e["sqrtA"]) < 2e-6
    r = np.linalg.norm(navsynth.sat_ecef(eq, 388800.0 + 100.0))
    assert 26.3e6 < r < 26.9e6


def test_scenario_geometry_and_frame_alignment():
    spec, truth = navsynth.build_scenario(seed=2)
    assert len(truth["prn"]) == 8 and len(set(truth["prn"])) == 8
    rng_ms = (np.array(truth["range"]) - min(truth["range"])) / navsynth.C * 1e3
    assert rng_ms.max() < 25.0                                   # GPS ranges differ by < 25 ms
    n = 38192
    for i, sat in enumerate(spec.sats):
        ns = truth["boundary_sample"][i]
        assert 5000 * n <= ns < 6000 * n                         # first boundary in [5000, 6000) ms
        # the code period starting at the boundary sample is the first of nav bit 300 (start of a subframe)
        cp = (int(spec.cp0[i]) + ns * int(spec.dcp[i]))
        assert cp % (1023 << 32) < int(spec.dcp[i])              # period boundary within one sample
        period = cp // (1023 << 32)
        assert (period + int(spec.per0[i])) % 20 == 0 and (period + int(spec.per0[i])) // 20 == 300
        assert abs(truth["doppler"][i]) < 5000.0
    # a few ms of signal around a boundary: the generator flips sign exactly where the stream says so
    assert synth.generate_cpu(spec, 4096).dtype == np.int8


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_decoder_recovers_the_ephemeris():
    from oracle import make_ref_shim
    warnings.filterwarnings("ignore")
    ref = make_ref_shim.import_ref()
    spec, truth = navsynth.build_scenario(seed=5)
    for i in (0, 3, 7):
        pm = spec.bits[i].astype(float)
        for w in range(1, 60):                                   # the reference's own parity check
            assert ref["postNavigation"].NavigationResult.navPartyChk(pm[30 * w - 2:30 * w + 30].copy()) != 0
        b01 = [str(int(x)) for x in (spec.bits[i] + 1) // 2]
        eph, tow = ref["ephemeris"].ephemeris(b01[300:1800], b01[299])
        assert tow == truth["tow"]
        names = ('weekNumber,accuracy,health,T_GD,IODC,t_oc,a_f2,a_f1,a_f0,IODE_sf2,C_rs,deltan,M_0,C_uc,e,C_us,'
                 'sqrtA,t_oe,C_ic,omega_0,C_is,i_0,C_rc,omega,omegaDot,IODE_sf3,iDot').split(',')
        d = dict(zip(names, eph))
        for k in ("M_0", "e", "sqrtA", "omega_0", "i_0", "omega", "omegaDot", "deltan", "t_oe"):
            assert d[k] == truth["eph"][i][k], k

"""BASELINE config 2 at full length on the GPU against the golden made by the REFERENCE's own caller
(tests/golden/make_golden_c2.py: ``Settings.postProcessing`` of initialize.py:454-527 run unmodified through the
Python-3 shim on the 37 100 ms LNAV recording of ``navsynth.build_scenario(seed=2)``):

  * the device generator reproduces the recording byte for byte (SHA-1);
  * acquisition (acquisition.py:27-204) and preRun (:259-306): PRN set, codePhase, channel table exact, carrFreq <= 1 Hz,
    peakMetric 1e-5 relative;
  * tracking of 8 channels x 37 000 ms (tracking.py:132-275), exact correlator variant, from the reference's channel
    table: ``absoluteSample`` identical at EVERY millisecond, sign(I_P) identical at every millisecond, and at every
    37th millisecond all 13 series: I/Q within 1e-9 of full scale, carrFreq 1e-7 Hz, codeFreq 1e-8 Hz, discriminators
    1e-9 (observed: 2e-11, 1 ulp);
  * the navigation chain on that tracking result (postNavigation.py:75-301): same measurement epochs, raw
    pseudoranges bit-identical, X/Y/Z/dt/height within 1e-5 m, latitude/longitude/az/el 1e-9 deg, DOP 1e-9.
"""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "c2_full.npz")
N = 38192


@pytest.fixture(scope="module")
def run():
    import torch
    from softgnss_python_b200 import _native, navsynth, synth
    from softgnss_python_b200.acquisition import AcquisitionResult
    from softgnss_python_b200.settings import Settings
    from softgnss_python_b200.tracking import TrackingResult
    g = np.load(GOLD, allow_pickle=False)
    ms, total = int(g["ms"]), int(g["total_ms"]) * N
    spec, truth = navsynth.build_scenario(seed=int(g["seed"]))
    L = _native.lib()
    stride = (total + 15) // 16 * 16
    dev = torch.empty((1, stride), dtype=torch.int8, device="cuda")
    sp, bits = _native.make_synth_specs([spec])
    L.synth(dev, stride, total, 0, sp, bits, synth.cos_lut(), _native.ca_chips_int8(), 0)
    torch.cuda.synchronize()
    sha = hashlib.sha1(dev[0, :total].cpu().numpy().tobytes()).hexdigest()
    s = Settings(msToProcess=float(ms), numberOfChannels=8)
    s.useTropCorr = False
    a = AcquisitionResult(s)
    a.acquire(dev[0, :11 * N])
    a.preRun()
    # tracking starts from the REFERENCE's channel table, so that this part pins tracking alone
    ref_ch = np.rec.fromarrays([g["ch_PRN"], g["ch_acquiredFreq"], g["ch_codePhase"], ['T'] * len(g["ch_PRN"])],
                               names="PRN,acquiredFreq,codePhase,status")

    class Acq(object):
        channels = ref_ch
        settings = s
    t = TrackingResult(Acq())
    t.track(dev[0, :total])
    return dict(g=g, sha=sha, acq=a, trk=t, settings=s, ms=ms)


def test_recording_is_reproduced_byte_for_byte(run):
    assert run["sha"] == str(run["g"]["recording_sha1"])


def test_acquisition_and_channel_table(run):
    g, a = run["g"], run["acq"]
    assert np.array_equal(a.carrFreq > 0, g["carrFreq"] > 0)
    assert np.array_equal(a.codePhase, g["codePhase"])
    assert np.abs(a.carrFreq - g["carrFreq"]).max() <= 1.0
    assert np.abs(a.peakMetric / g["peakMetric"] - 1).max() <= 1e-5
    assert np.array_equal(a.channels.PRN, g["ch_PRN"])
    assert np.array_equal(a.channels.codePhase, g["ch_codePhase"])
    assert np.abs(a.channels.acquiredFreq - g["ch_acquiredFreq"]).max() <= 1.0


def test_tracking_every_millisecond(run):
    g, r, ms = run["g"], run["trk"].results, run["ms"]
    assert np.array_equal(np.asarray(r.PRN), g["trk_PRN"])
    start = run["settings"].skipNumberOfBytes + g["ch_codePhase"][:len(r)]
    ref_abs = start[:, None] + np.cumsum(g["abs_delta"].astype(np.float64) + N, axis=1)
    got_abs = np.stack([np.asarray(x, dtype=np.float64) for x in r.absoluteSample])
    assert got_abs.shape == (8, ms)
    assert np.array_equal(got_abs, ref_abs), "absoluteSample differs at %d of %d ms" % (
        int((got_abs != ref_abs).sum()), got_abs.size)
    ip = np.stack([np.asarray(x, dtype=np.float64) for x in r.I_P])
    assert np.array_equal(np.packbits(ip > 0, axis=1), g["ip_sign"]), "sign(I_P)"
    assert np.array_equal(np.argwhere(ip == 0), g["ip_zero"].reshape(-1, 2))


def test_tracking_series_every_37th_millisecond(run):
    from softgnss_python_b200._native import TRACK_FIELDS
    g, r = run["g"], run["trk"].results
    sub = int(g["sub"])
    ref = g["sub_series"]                                            # [8, 13, ms / 37]
    got = np.stack([np.stack([np.asarray(r[c][f], dtype=np.float64)[::sub] for f in TRACK_FIELDS]) for c in range(len(r))])
    assert got.shape == ref.shape
    scale = np.abs(ref[:, 3:9]).max()
    tol = dict(absoluteSample=0.0, codeFreq=1e-8, carrFreq=1e-7, dllDiscr=1e-9, dllDiscrFilt=1e-8, pllDiscr=1e-9,
               pllDiscrFilt=1e-7)
    for i, f in enumerate(TRACK_FIELDS):
        t = tol.get(f, 1e-9 * scale)
        err = np.abs(got[:, i] - ref[:, i]).max()
        assert err <= t, "%s: max deviation %.3g > %.3g" % (f, err, t)


def test_navigation_solutions(run):
    from softgnss_python_b200 import postnav
    from tests.nav_util import ANGLE_DEG, DOP_ABS, POS_M
    g = run["g"]
    nav, eph = postnav.postNavigate(run["trk"].results, run["settings"])
    assert nav is not None
    sol = nav[0]
    n = int(np.sum(~np.isnan(g["sol_X"])))
    assert n > 50 and len(sol.X) >= n

    def close(got, want, tol, name):
        got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
        assert np.array_equal(np.isnan(got), np.isnan(want)), name + ": NaN pattern"
        fin = np.isfinite(want)
        err = np.abs(got[fin] - want[fin]).max() if fin.any() else 0.0
        assert err <= tol, "%s: max error %.3g > %.3g" % (name, err, tol)
    for f in ("X", "Y", "Z", "dt", "height"):
        close(getattr(sol, f)[:n], g["sol_" + f][:n], POS_M, f)
    for f in ("latitude", "longitude"):
        close(getattr(sol, f)[:n], g["sol_" + f][:n], ANGLE_DEG, f)
    close(np.asarray(sol.DOP)[:, :n], g["sol_DOP"][:, :n], DOP_ABS, "DOP")
    ch = sol.channel[0]
    assert np.array_equal(np.asarray(ch.rawP)[:, :n], g["solch_rawP"][:, :n], equal_nan=True), "raw pseudoranges"
    close(np.asarray(ch.correctedP)[:, :n], g["solch_correctedP"][:, :n], POS_M, "correctedP")
    close(np.asarray(ch.el)[:, :n], g["solch_el"][:, :n], ANGLE_DEG, "el")
    close(np.asarray(ch.az)[:, :n], g["solch_az"][:, :n], ANGLE_DEG, "az")
    assert np.array_equal(np.asarray(ch.PRN)[:, :n], g["solch_PRN"][:, :n]), "satellites used per epoch"

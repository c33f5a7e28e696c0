"""sgx_nav_solve on the B200 (second half): the measurement loop of postNavigate
against the reference's golden output (tests/golden/nav.npz), the oracle, batch invariance, device-resident
tracking output, and the reference-layout wrapper."""
import numpy as np
import pytest

from oracle import gnss_oracle as orc
from tests import nav_util
from tests.cases import NAV_MS, load_nav_cases, nav_abs_sample

pytestmark = pytest.mark.gpu


def _run(case, abs_sample=None, **kw):
    from softgnss_python_b200 import postnav
    n_ch = len(case["prn"])
    s = nav_util.settings_for(case, n_ch, NAV_MS)
    sfs, ready, eph = nav_util.case_inputs(case)
    if abs_sample is None:
        abs_sample = nav_abs_sample(case["coef"])[None]
    return postnav.nav_solve_batch(abs_sample, sfs[None], ready[None], eph[None], [case["tow"]], s, **kw)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_nav_solve_matches_reference_golden(idx):
    case = load_nav_cases()[idx]
    out = _run(case)
    worst = nav_util.compare_nav(out, 0, case["ref"], "case %d" % idx)
    print("case %d worst errors vs the reference:" % idx, {k: float("%.2g" % v) for k, v in worst.items()})


def test_nav_solve_matches_oracle_with_other_settings():
    # settings the golden does not cover: every case under the other cases' mask / correction switch
    cases = load_nav_cases()
    for idx, case in enumerate(cases):
        for other in cases:
            c = dict(case, elevation_mask=other["elevation_mask"], use_trop_corr=not case["use_trop_corr"])
            o = orc.nav_solve(nav_abs_sample(c["coef"]), c["prn"], c["sub_frame_start"], c["ready"], c["eph"], c["tow"],
                              float(NAV_MS), 38192, elevation_mask=c["elevation_mask"], use_trop_corr=c["use_trop_corr"])
            nav_util.compare_nav(_run(c), 0, o, "case %d mask %.1f" % (idx, c["elevation_mask"]))


def test_nav_solve_batch_invariance_and_device_input():
    import torch
    from softgnss_python_b200 import postnav
    cases = load_nav_cases()
    case = cases[0]
    n_ch = len(case["prn"])
    s = nav_util.settings_for(case, n_ch, NAV_MS)
    # 37 recordings: the three golden geometries repeated, each shifted by a whole number of code periods
    R = 37
    abs_all, sfs_all, ready_all, eph_all, tow_all = [], [], [], [], []
    for r in range(R):
        c = cases[r % 3]
        sfs, ready, eph = nav_util.case_inputs(c)
        abs_all.append(nav_abs_sample(c["coef"]) + 38192.0 * (r // 3))
        sfs_all.append(sfs); ready_all.append(ready); eph_all.append(eph); tow_all.append(c["tow"])
    abs_all = np.stack(abs_all)
    out = postnav.nav_solve_batch(abs_all, np.stack(sfs_all), np.stack(ready_all), np.stack(eph_all), tow_all, s)
    for r in range(R):
        c = dict(cases[r % 3], elevation_mask=s.elevationMask, use_trop_corr=s.useTropCorr)
        single = _run(c, abs_sample=abs_all[r:r + 1])
        for k in ("rawP", "correctedP", "el", "az", "satPositions", "satClkCorr", "active", "sol"):
            assert np.array_equal(out[k][r], single[k][0], equal_nan=True), (r, k)     # batch position does not matter
    # a whole-period shift of every channel leaves the relative pseudoranges, hence the fix, unchanged
    assert np.array_equal(out["sol"][0], out["sol"][3], equal_nan=True)
    # tracking-output layout [R, C, 13, ms] resident on the device (field 0 read in place)
    trk = torch.zeros((R, n_ch, 13, NAV_MS), dtype=torch.float64, device="cuda")
    trk[:, :, 0, :] = torch.from_numpy(abs_all).cuda()
    dev = postnav.nav_solve_batch(trk, np.stack(sfs_all), np.stack(ready_all), np.stack(eph_all), tow_all, s)
    for k in ("rawP", "sol", "el", "active"):
        assert np.array_equal(dev[k], out[k], equal_nan=True), k


def test_navSolutions_reference_layout():
    from softgnss_python_b200 import postnav
    case = load_nav_cases()[1]
    n_ch = len(case["prn"])
    s = nav_util.settings_for(case, n_ch, NAV_MS)
    abs_sample = nav_abs_sample(case["coef"])

    class Rec(object):
        pass
    trk = []
    for c in range(n_ch):
        r = Rec(); r.absoluteSample = abs_sample[c]; r.PRN = int(case["prn"][c]); trk.append(r)
    nav, channel = postnav.navSolutions(trk, case["sub_frame_start"], case["ready"], case["eph"], case["tow"], s)
    ref = case["ref"]
    assert np.array_equal(channel[0].PRN, ref["PRN"])
    assert np.array_equal(channel[0].rawP, ref["rawP"], equal_nan=True)
    assert np.abs(nav[0].X - ref["X"]).max() < nav_util.POS_M and nav[0].DOP.shape == ref["DOP"].shape
    assert np.abs(nav[0].latitude - ref["latitude"]).max() < nav_util.ANGLE_DEG


def test_nav_solve_argument_errors():
    from softgnss_python_b200 import _native, postnav
    case = load_nav_cases()[0]
    n_ch = len(case["prn"])
    s = nav_util.settings_for(case, n_ch, NAV_MS)
    sfs, ready, eph = nav_util.case_inputs(case)
    with pytest.raises(_native.NativeError):
        _native.lib().nav_solve(np.zeros((1, 33, 10)), 10, 1, 33, 10, np.zeros((1, 33)), np.zeros((1, 33)),
                                np.zeros((1, 33, 21)), [0.0], [1], postnav.nav_settings(s))


def test_post_navigate_chain_matches_oracle_chain():
    """Preamble search -> ephemeris decoding -> measurement loop (postnav.post_navigate_batch) on synthetic tracking
    results (LNAV streams in I_P, geometry in absoluteSample), host and device resident, against the oracle's chain;
    the fix is the scenario's antenna."""
    import torch
    from softgnss_python_b200 import postnav
    from softgnss_python_b200.settings import Settings
    from tests.cases import build_chain_case, oracle_chain
    tr0, prn0, truth0 = build_chain_case(seed=2)
    tr1, prn1, truth1 = build_chain_case(seed=5, drop=(1, 4, 6, 7, 2))       # three satellites left: no solution
    tr2, prn2, truth2 = build_chain_case(seed=7, drop=(3,))
    s = Settings(numberOfChannels=8, msToProcess=float(NAV_MS))
    kw = dict(elevation_mask=s.elevationMask, use_trop_corr=s.useTropCorr)
    track = np.stack([tr0, tr1, tr2])
    prn = np.stack([prn0, prn1, prn2])
    out = postnav.post_navigate_batch(track, prn, s)
    dev = postnav.post_navigate_batch(torch.from_numpy(track).cuda(), prn, s)
    for k in ("sol", "rawP", "el", "active", "subFrameStart", "ready", "n_epochs"):
        assert np.array_equal(out[k], dev[k], equal_nan=True), k              # device-resident input: same result
    for r, (tr, p, truth) in enumerate(((tr0, prn0, truth0), (tr1, prn1, truth1), (tr2, prn2, truth2))):
        first, o = oracle_chain(tr, p, kw)
        assert np.array_equal(out["subFrameStart"][r], first)
        if o is None:
            assert out["n_epochs"][r] == 0
            continue
        nav_util.compare_nav(out, r, o, "chain %d" % r)
        sol = out["sol"][r, :out["n_epochs"][r]]
        err = np.sqrt(((sol[:, :3] - truth["rx"]) ** 2).sum(1))
        assert np.isfinite(err).all() and np.median(err) < 30.0, np.median(err)
    assert out["n_epochs"].tolist() == [63, 0, 63]


def test_postNavigate_reference_surface(capsys):
    from softgnss_python_b200 import postnav
    from softgnss_python_b200.settings import Settings
    from tests.cases import build_chain_case
    tr, prn, truth = build_chain_case(seed=2)
    s = Settings(numberOfChannels=8, msToProcess=float(NAV_MS))
    dtype = [('status', 'U1')] + [(f, 'object') for f in orc.TRACK_FIELDS] + [('PRN', 'int64')]
    res = np.rec.fromrecords([('T',) + tuple(tr[c, i] for i in range(13)) + (int(prn[c]),) for c in range(8)], dtype=dtype)
    nav, eph = postnav.postNavigate(res, s)
    assert nav[0].X.shape == (63,) and np.isfinite(nav[0].X).all()
    err = np.sqrt((nav[0].X - truth["rx"][0]) ** 2 + (nav[0].Y - truth["rx"][1]) ** 2 + (nav[0].Z - truth["rx"][2]) ** 2)
    assert np.median(err) < 30.0
    assert abs(eph[int(prn[0]) - 1]["sqrtA"] - truth["eph"][0]["sqrtA"]) < 1e-9
    s2 = Settings(numberOfChannels=8, msToProcess=20000.0)
    assert postnav.postNavigate(res, s2) == (None, None)                       # postNavigation.py:104-111
    assert "Record is to short" in capsys.readouterr().out

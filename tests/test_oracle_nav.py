"""The oracle's navigation-solution restatement (oracle/gnss_oracle.py: satpos, least_square_pos, nav_solve)
against the reference's own measurement loop (tests/golden/nav.npz, made by tests/golden/make_golden_nav.py
through the Python-3 shim).  SURVEY.md section 8(f) row 4."""
import numpy as np
import pytest

from oracle import gnss_oracle as orc
from tests.cases import NAV_MS, load_nav_cases, nav_abs_sample

KEYS = ("rawP", "el", "az", "correctedP", "DOP", "X", "Y", "Z", "dt", "latitude", "longitude", "height", "PRN")


def run_oracle(case):
    abs_sample = nav_abs_sample(case["coef"])
    return orc.nav_solve(abs_sample, case["prn"], case["sub_frame_start"], case["ready"], case["eph"], case["tow"],
                         float(NAV_MS), 38192, elevation_mask=case["elevation_mask"], use_trop_corr=case["use_trop_corr"])


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_nav_solve_matches_reference(idx):
    case = load_nav_cases()[idx]
    got = run_oracle(case)
    exact = True
    for k in KEYS:
        ref = case["ref"][k]
        assert got[k].shape == ref.shape, k
        assert np.array_equal(np.isnan(got[k]), np.isnan(ref)), k           # same "no fix" / masked pattern
        assert np.array_equal(np.isinf(got[k]), np.isinf(ref)), k           # infinite pseudoranges of unlisted channels
        fin = np.isfinite(ref)
        # float64 tolerance: 1e-12 relative (same numpy calls in the same order; normally bit-identical)
        np.testing.assert_allclose(got[k][fin], ref[fin], rtol=1e-12, atol=0, err_msg=k)
        exact = exact and np.array_equal(got[k][fin], ref[fin])
    if idx == 0:
        assert np.isfinite(got["X"]).all()
        err = np.sqrt((got["X"] - case["rx"][0]) ** 2 + (got["Y"] - case["rx"][1]) ** 2 + (got["Z"] - case["rx"][2]) ** 2)
        assert np.median(err) < 50.0                                          # the fix is the scenario's receiver
    if idx == 1:
        assert (got["PRN"][:, 0] > 0).sum() == 8 and (got["PRN"][:, 1] > 0).sum() == 6   # elevation mask applied
    if idx == 2:
        assert np.isfinite(got["X"]).sum() == 1 and (got["DOP"][:, 1:] == 0).all()        # "not enough information"
    print("case %d bit-identical to the reference: %s" % (idx, exact))


def test_tropo_and_topocent_known_values():
    # zenith delay of the standard atmosphere used by leastSquarePos (:697) is about 2.4 m and grows towards the horizon
    z = orc.tropo(1.0, 0.0, 1013.0, 293.0, 50.0, 0.0, 0.0, 0.0)
    h = orc.tropo(np.sin(np.radians(10.0)), 0.0, 1013.0, 293.0, 50.0, 0.0, 0.0, 0.0)
    assert 2.3 < z < 2.6 and 12.0 < h < 15.0
    az, el, d = orc.topocent(np.array([6378137.0, 0.0, 0.0]), np.array([100.0, 0.0, 0.0]))
    assert abs(el - 90.0) < 1e-6 and abs(d - 100.0) < 1e-9
    az, el, d = orc.topocent(np.array([6378137.0, 0.0, 0.0]), np.array([0.0, 0.0, 100.0]))
    assert abs(az) < 1e-9 and abs(el) < 1e-6                                  # due north on the horizon


def test_ephemeris_decoder_matches_reference():
    """oracle.ephemeris and the product's host-side decoder (postnav.ephemeris) against ephemeris.py of the
    reference (tests/golden/ephemeris.npz): every field and the TOW identical on encoder streams and random bits."""
    import hashlib
    import os
    from softgnss_python_b200 import postnav
    from tests.cases import build_ephemeris_cases
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ephemeris.npz"))
    rows = build_ephemeris_cases()
    assert hashlib.sha1(rows.tobytes()).hexdigest() == str(g["input_sha1"])
    for r, row in enumerate(rows):
        for dec in (orc.ephemeris, postnav.ephemeris):
            eph, tow = dec(row[1:], row[0])
            assert tow == g["table"][r, 27]
            assert [eph[k] for k in orc.EPH_ALL] == g["table"][r, :27].tolist(), (r, dec.__module__)
    # a stream without subframes 1-3 leaves their fields undecoded
    eph, _ = postnav.ephemeris(np.zeros(1500, dtype=np.uint8), 0)
    assert eph["IODC"] is None and eph["IODE_sf2"] is None

"""Golden vectors for the navigation solution (second half), produced by the
REFERENCE's own measurement loop -- ``NavigationResult.postNavigate`` (postNavigation.py:159-301) with its
``calculatePseudoranges``, ``geoFunctions.satpos``, ``leastSquarePos`` (``e_r_corr``, ``topocent``, ``togeod``,
``tropo``) and ``cart2geo`` -- through the Python-3 shim.  Build container only:

    python tests/golden/make_golden_nav.py

The stages in front of the loop (preamble search, ephemeris decoding) are replaced by stubs that return the
case's subframe starts / ephemerides, so that the loop can be driven without a 37 s recording.  The float
inputs are stored next to the outputs (tests/cases.py:load_nav_cases rebuilds the absoluteSample series from
integer coefficients)."""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import make_ref_shim                                  # noqa: E402
from oracle.gnss_oracle import TRACK_FIELDS                      # noqa: E402
from softgnss_python_b200 import navsynth                        # noqa: E402
from tests.cases import NAV_FIELDS, NAV_MS, nav_abs_sample       # noqa: E402

FS = 38.192e6
N_CODE = 38192
EPH_NAMES = ('weekNumber,accuracy,health,T_GD,IODC,t_oc,a_f2,a_f1,a_f0,IODE_sf2,C_rs,deltan,M_0,C_uc,e,C_us,sqrtA,'
             't_oe,C_ic,omega_0,C_is,i_0,C_rc,omega,omegaDot,IODE_sf3,iDot').split(',')


def scenario_inputs(seed, first_boundary_ms=5200):
    """Ephemerides with clock terms + integer arrival polynomials from the geometry of navsynth.build_scenario."""
    _, truth = navsynth.build_scenario(seed=seed)
    rng = np.random.default_rng(1000 + seed)
    n_ch = len(truth["prn"])
    eph_arr = np.zeros((n_ch, len(NAV_FIELDS)))
    coef = np.zeros((n_ch, 4), dtype=np.int64)
    rho_min = min(truth["range"])
    elev = []
    for c in range(n_ch):
        e = dict(truth["eph"][c])
        e.update(t_oc=e["t_oe"], a_f2=0.0, a_f1=float(rng.integers(-200, 200)) * 2.0 ** -43,
                 a_f0=float(rng.integers(-2 ** 18, 2 ** 18)) * 2.0 ** -31, T_GD=float(rng.integers(-20, 20)) * 2.0 ** -31)
        eph_arr[c] = [e[k] for k in NAV_FIELDS]
        k0 = first_boundary_ms + int(rng.integers(0, 4))
        ks = np.array([0.0, 18000.0, 36000.0])
        s = []
        for k in ks:                        # arrival sample (relative to k * N_CODE) of the code period sent at tow + k ms
            rho, xs = navsynth.geometric_range(e, truth["tow"] + k * 1e-3, truth["rx"])
            s.append(((rho - rho_min) / navsynth.C - e["a_f0"]) * FS)
            if k == 0:
                elev.append(navsynth.elevation(truth["rx"], xs))
        p2 = ((s[2] - s[0]) - 2 * (s[1] - s[0])) / (2 * 18000.0 ** 2)
        p1 = (s[1] - s[0]) / 18000.0 - p2 * 18000.0
        coef[c] = (first_boundary_ms * N_CODE + int(round(s[0])), k0, int(round(p1 * 2.0 ** 40)), int(round(p2 * 2.0 ** 40)))
    return truth, eph_arr, coef, np.array(elev)


def run_reference(ref, case):
    s = ref["initialize"].Settings()
    s.msToProcess = float(NAV_MS)
    n_ch = len(case["prn"])
    s.numberOfChannels = n_ch
    s.elevationMask = case["elevation_mask"]
    s.useTropCorr = case["use_trop_corr"]
    abs_sample = nav_abs_sample(case["coef"], N_CODE, NAV_MS)
    dtype = [('status', 'S1')] + [(f, 'object') for f in TRACK_FIELDS] + [('PRN', 'int64')]
    zero = np.zeros(NAV_MS)
    rec = [(b'T',) + tuple(abs_sample[c] if f == "absoluteSample" else zero for f in TRACK_FIELDS) + (int(case["prn"][c]),)
           for c in range(n_ch)]
    res = np.rec.fromrecords(rec, dtype=dtype)

    class T(object):
        results = res
        channels = None
        settings = s
    nav = ref["postNavigation"].NavigationResult(T())
    nav.findPreambles = lambda: (case["sub_frame_start"].copy(), case["ready"].copy())     # stage stub
    order = iter(case["ready"].tolist())

    def fake_ephemeris(bits, d30star):                                                      # stage stub
        c = next(order)
        e = dict(zip(NAV_FIELDS, case["eph_arr"][c]))
        e.update(weekNumber=1076, accuracy=0, health=0, IODC=1, IODE_sf2=1, IODE_sf3=1)
        return tuple(e[k] for k in EPH_NAMES), case["tow"]
    real_ephemeris = ref["postNavigation"].ephemeris.ephemeris
    ref["postNavigation"].ephemeris.ephemeris = fake_ephemeris
    try:
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            nav.postNavigate()
    finally:
        ref["postNavigation"].ephemeris.ephemeris = real_ephemeris
    sol = nav.solutions[0]
    n_ep = int(np.fix(NAV_MS - case["sub_frame_start"].max()) / s.navSolPeriod)
    out = {k: np.array(getattr(sol.channel[0], k)[:, :n_ep], dtype=np.float64) for k in ("rawP", "el", "az", "correctedP", "PRN")}
    out["DOP"] = np.array(sol.DOP[:, :n_ep], dtype=np.float64)
    for k in ("X", "Y", "Z", "dt", "latitude", "longitude", "height"):
        out[k] = np.array(getattr(sol, k)[:n_ep], dtype=np.float64)
    return out


def main():
    ref = make_ref_shim.import_ref()
    store = {}
    cases = []
    # case 0: 8 satellites, reference defaults (mask 10 deg, tropospheric correction on)
    # case 1: mask raised so that the lowest satellites drop out after the first fix; correction off
    # case 2: 5 of 8 channels have an ephemeris and the mask removes two of them -> "not enough information"
    for idx, seed in enumerate((2, 5, 7)):
        truth, eph_arr, coef, elev = scenario_inputs(seed)
        n_ch = len(truth["prn"])
        sfs = coef[:, 1].copy()
        ready = np.arange(n_ch)
        mask, trop = 10.0, True
        order = np.argsort(elev)
        if idx == 1:
            mask, trop = float(0.5 * (elev[order[1]] + elev[order[2]])), False      # two satellites below the mask
        if idx == 2:
            ready = np.sort(order[[0, 1, 4, 6, 7]])                                   # two low + three high satellites
            mask = float(0.5 * (elev[order[1]] + elev[order[4]]))
            sfs[np.setdiff1d(np.arange(n_ch), ready)] = 0                             # no preamble found there
        case = dict(coef=coef, prn=np.array(truth["prn"], dtype=np.int64), eph_arr=eph_arr, sub_frame_start=sfs,
                    ready=ready, tow=float(truth["tow"]), elevation_mask=mask, use_trop_corr=trop, rx=truth["rx"])
        out = run_reference(ref, case)
        cases.append((case, out))
        for k, v in case.items():
            store["c%d_%s" % (idx, k)] = np.asarray(v)
        for k, v in out.items():
            store["c%d_ref_%s" % (idx, k)] = v
        err = np.sqrt((out["X"] - truth["rx"][0]) ** 2 + (out["Y"] - truth["rx"][1]) ** 2 + (out["Z"] - truth["rx"][2]) ** 2)
        print("case %d: epochs %d, elevations %s, mask %.2f, fixes %d, |pos - truth| median %.1f m, active per epoch %s"
              % (idx, out["X"].size, np.round(np.sort(elev), 1), mask, np.isfinite(out["X"]).sum(),
                 np.nanmedian(err) if np.isfinite(err).any() else np.nan, (out["PRN"] > 0).sum(0)[:4]))
    store["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "nav.npz"), **store)
    make_ephemeris_golden(ref)


def make_ephemeris_golden(ref):
    """ephemeris.py:98-196 of the reference on the streams of tests/cases.py:build_ephemeris_cases ->
    tests/golden/ephemeris.npz (27 decoded fields in the reference's order + TOW per stream)."""
    import hashlib
    from tests.cases import build_ephemeris_cases
    rows = build_ephemeris_cases()
    table = np.zeros((len(rows), 28))
    for r, row in enumerate(rows):
        chars = [str(int(b)) for b in row]
        eph, tow = ref["ephemeris"].ephemeris(chars[1:], chars[0])
        table[r, :27] = eph
        table[r, 27] = tow
    np.savez_compressed(os.path.join(HERE, "ephemeris.npz"), table=table,
                        input_sha1=hashlib.sha1(rows.tobytes()).hexdigest())
    print("ephemeris: %d streams, TOW %s" % (len(rows), table[:, 27].astype(int).tolist()))


if __name__ == "__main__":
    main()

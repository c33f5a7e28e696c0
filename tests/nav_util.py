"""Comparison of the device navigation solution (sgx_nav_solve) with the reference / oracle, shared by the -m gpu
tests and the developer's CPU-emulator check."""
import numpy as np

from softgnss_python_b200 import _native

# Tolerances (float64 on both sides; the device differs from numpy in the last-ulp rounding of sin/cos/atan2/pow
# and in the solver of the 8x4 least-squares step, Cholesky on the normal equations instead of LAPACK gelsd):
POS_M = 1e-5        # X, Y, Z, dt, height, corrected pseudoranges: metres (observed: 1e-7)
ANGLE_DEG = 1e-9    # el, az, latitude, longitude: degrees (1e-9 deg = 0.1 mm on the ground)
DOP_ABS = 1e-9      # dilution-of-precision values (order 1..5), absolute
SAT_M = 1e-6        # satellite ECEF coordinates, metres (observed: 2e-8);  clock corrections: 1e-15 s


class _S(object):
    pass


def settings_for(case, n_ch, ms):
    s = _S()
    s.samplesPerCode, s.startOffset, s.c, s.navSolPeriod = 38192, 68.802, 299792458.0, 500.0
    s.elevationMask, s.useTropCorr = case["elevation_mask"], case["use_trop_corr"]
    s.msToProcess, s.numberOfChannels = float(ms), n_ch
    return s


def case_inputs(case):
    n_ch = len(case["prn"])
    ready = np.zeros(n_ch, dtype=np.uint8)
    ready[case["ready"]] = 1
    return case["sub_frame_start"].astype(np.int32), ready, case["eph_arr"]


def compare_nav(out, r, ref, label=""):
    """out: dict from Lib.nav_solve; r: recording index; ref: dict in the reference's layout ([C, E] / [E] / [5, E])."""
    n_ep = ref["X"].size
    assert out["n_epochs"][r] == n_ep, label
    sol = out["sol"][r, :n_ep]
    act = out["active"][r, :n_ep].T > 0
    assert np.array_equal(act, ref["PRN"] > 0), label + " channel lists"
    assert np.array_equal(out["rawP"][r, :n_ep].T, ref["rawP"], equal_nan=True), label + " rawP (bit-identical)"

    def close(got, want, tol, name):
        assert np.array_equal(np.isnan(got), np.isnan(want)), "%s %s: NaN pattern" % (label, name)
        fin = np.isfinite(want)
        err = np.abs(got[fin] - want[fin])
        assert err.size == 0 or err.max() <= tol, "%s %s: max error %.3g > %.3g" % (label, name, err.max(), tol)
        return err.max() if err.size else 0.0
    worst = {}
    for k, name in enumerate(_native.NAV_SOL_FIELDS):
        if name in ("X", "Y", "Z", "dt", "height"):
            worst[name] = close(sol[:, k], ref[name], POS_M, name)
        elif name in ("latitude", "longitude"):
            worst[name] = close(sol[:, k], ref[name], ANGLE_DEG, name)
    worst["DOP"] = close(sol[:, 4:9].T, ref["DOP"], DOP_ABS, "DOP")
    worst["el"] = close(out["el"][r, :n_ep].T, ref["el"], ANGLE_DEG, "el")
    worst["az"] = close(out["az"][r, :n_ep].T, ref["az"], ANGLE_DEG, "az")
    worst["correctedP"] = close(out["correctedP"][r, :n_ep].T, ref["correctedP"], POS_M, "correctedP")
    if "satPositions" in ref and out.get("satPositions") is not None:
        worst["sat"] = close(out["satPositions"][r, :n_ep].transpose(1, 0, 2), ref["satPositions"], SAT_M, "satPositions")
        worst["clk"] = close(out["satClkCorr"][r, :n_ep].T, ref["satClkCorr"], 1e-15, "satClkCorr")
    return worst

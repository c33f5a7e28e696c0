"""Synthetic GPS LNAV navigation data and geometry for full-pipeline recordings (BASELINE config 2/4).

The reference's downstream stage (``postNavigation.py``) needs, per tracked channel, a 50 bit/s
stream with the TLM preamble every 6 s, words that pass its parity check
(``postNavigation.py:443-521``) and subframes 1-3 carrying an ephemeris in the layout its decoder
slices (``ephemeris.py:98-190``); and it needs the subframe boundaries of the satellites to arrive
staggered by their geometric ranges so that ``leastSquarePos`` converges.  This module is the
*encoder* for that wire format plus a Kepler propagator to make the delays consistent:

* :func:`lnav_word` / :func:`encode_subframe` -- IS-GPS-200 word parity (D25..D30, D29*/D30* chaining,
  data inversion by D30*) and the field layout of subframes 1, 2, 3 (4 and 5: TLM/HOW only);
* :func:`sat_ecef` -- broadcast-ephemeris orbit (same model the receiver applies in
  ``geoFunctions.satpos``, written from IS-GPS-200 table 20-IV);
* :func:`build_scenario` -- receiver position + 8 satellites -> ``synth.SatSpec`` list whose code
  phase, nav-bit alignment and Doppler follow from the geometry, and the truth needed to check a
  position fix.

Everything here is host-side test/bench input generation.
"""
import numpy as np

from . import synth

GPS_PI = 3.1415926535898
GM = 3.986005e14
OMEGA_E = 7.2921151467e-5
C = 299792458.0
PREAMBLE = (1, 0, 0, 0, 1, 0, 1, 1)

# parity equations of IS-GPS-200 table 20-XIV: data-bit numbers (1-based) entering D25..D30
_PAR = (
    (1, 2, 3, 5, 6, 10, 11, 12, 13, 14, 17, 18, 20, 23),
    (2, 3, 4, 6, 7, 11, 12, 13, 14, 15, 18, 19, 21, 24),
    (1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22),
    (2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23),
    (1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24),
    (3, 5, 6, 8, 9, 10, 11, 13, 15, 19, 22, 23, 24),
)
_PAR_STAR = (0, 1, 0, 1, 1, 0)   # 0: D29*, 1: D30*


def lnav_word(d, d29s, d30s):
    """24 source data bits -> 30 transmitted bits (data XOR D30*, then D25..D30)."""
    d = [int(b) for b in d]
    assert len(d) == 24
    out = [b ^ d30s for b in d]
    for eq, star in zip(_PAR, _PAR_STAR):
        p = d30s if star else d29s
        for k in eq:
            p ^= d[k - 1]
        out.append(p)
    return out


def _ubits(value, n):
    value = int(value)
    assert 0 <= value < (1 << n), (value, n)
    return [(value >> (n - 1 - i)) & 1 for i in range(n)]


def _sbits(value, n):
    value = int(value)
    assert -(1 << (n - 1)) <= value < (1 << (n - 1)), (value, n)
    return _ubits(value & ((1 << n) - 1), n)


def quantize_ephemeris(e):
    """Round orbital elements to the LNAV scale factors (the values the receiver will decode)."""
    q = {}
    q["M_0"] = int(round(e["M_0"] / GPS_PI * 2 ** 31))
    q["deltan"] = int(round(e.get("deltan", 0.0) / GPS_PI * 2 ** 43))
    q["e"] = int(round(e["e"] * 2 ** 33))
    q["sqrtA"] = int(round(e["sqrtA"] * 2 ** 19))
    q["omega_0"] = int(round(e["omega_0"] / GPS_PI * 2 ** 31))
    q["i_0"] = int(round(e["i_0"] / GPS_PI * 2 ** 31))
    q["omega"] = int(round(e["omega"] / GPS_PI * 2 ** 31))
    q["omegaDot"] = int(round(e.get("omegaDot", 0.0) / GPS_PI * 2 ** 43))
    q["iDot"] = int(round(e.get("iDot", 0.0) / GPS_PI * 2 ** 43))
    q["t_oe"] = int(round(e["t_oe"] / 16))
    q["t_oc"] = q["t_oe"]
    q["weekNumber"] = int(e.get("weekNumber", 2100)) % 1024
    for k in ("C_rs", "C_rc"):
        q[k] = int(round(e.get(k, 0.0) * 2 ** 5))
    for k in ("C_uc", "C_us", "C_ic", "C_is"):
        q[k] = int(round(e.get(k, 0.0) * 2 ** 29))
    q["IODE"] = int(e.get("IODE", 77)) & 0xFF
    return q


def dequantize_ephemeris(q):
    """What ``ephemeris.py`` recovers from the encoded fields (radians, metres, seconds)."""
    return dict(M_0=q["M_0"] * 2.0 ** -31 * GPS_PI, deltan=q["deltan"] * 2.0 ** -43 * GPS_PI,
                e=q["e"] * 2.0 ** -33, sqrtA=q["sqrtA"] * 2.0 ** -19,
                omega_0=q["omega_0"] * 2.0 ** -31 * GPS_PI, i_0=q["i_0"] * 2.0 ** -31 * GPS_PI,
                omega=q["omega"] * 2.0 ** -31 * GPS_PI, omegaDot=q["omegaDot"] * 2.0 ** -43 * GPS_PI,
                iDot=q["iDot"] * 2.0 ** -43 * GPS_PI, t_oe=q["t_oe"] * 16.0,
                C_rs=q["C_rs"] * 2.0 ** -5, C_rc=q["C_rc"] * 2.0 ** -5, C_uc=q["C_uc"] * 2.0 ** -29,
                C_us=q["C_us"] * 2.0 ** -29, C_ic=q["C_ic"] * 2.0 ** -29, C_is=q["C_is"] * 2.0 ** -29)


def _subframe_data(q, sf_id, tow_count):
    """10 x 24 source data bits of one subframe; positions follow the slices in ephemeris.py:117-173
    (index i of the 300-bit subframe = word i//30, bit i%30)."""
    bits = [0] * 300

    def put(lo, hi, field):
        assert hi - lo == len(field)
        bits[lo:hi] = field

    put(0, 8, list(PREAMBLE))                       # TLM preamble (postNavigation.py:556)
    put(30, 47, _ubits(tow_count, 17))              # HOW: TOW count (ephemeris.py:190)
    put(49, 52, _ubits(sf_id, 3))                   # HOW: subframe ID (ephemeris.py:110)
    if sf_id == 1:
        put(60, 70, _ubits(q["weekNumber"], 10))
        put(72, 76, _ubits(0, 4))                   # accuracy
        put(76, 82, _ubits(0, 6))                   # health
        put(82, 84, _ubits(0, 2))                   # IODC msbs
        put(218, 234, _ubits(q["t_oc"], 16))
        put(240, 248, _sbits(0, 8))                 # a_f2
        put(248, 264, _sbits(0, 16))                # a_f1
        put(270, 292, _sbits(0, 22))                # a_f0   (T_GD at 195:204 stays zero)
    elif sf_id == 2:
        put(60, 68, _ubits(q["IODE"], 8))
        put(68, 84, _sbits(q["C_rs"], 16))
        put(90, 106, _sbits(q["deltan"], 16))
        m0 = _sbits(q["M_0"], 32)
        put(106, 114, m0[:8]); put(120, 144, m0[8:])
        put(150, 166, _sbits(q["C_uc"], 16))
        ee = _ubits(q["e"], 32)
        put(166, 174, ee[:8]); put(180, 204, ee[8:])
        put(210, 226, _sbits(q["C_us"], 16))
        sa = _ubits(q["sqrtA"], 32)
        put(226, 234, sa[:8]); put(240, 264, sa[8:])
        put(270, 286, _ubits(q["t_oe"], 16))
    elif sf_id == 3:
        put(60, 76, _sbits(q["C_ic"], 16))
        o0 = _sbits(q["omega_0"], 32)
        put(76, 84, o0[:8]); put(90, 114, o0[8:])
        put(120, 136, _sbits(q["C_is"], 16))
        i0 = _sbits(q["i_0"], 32)
        put(136, 144, i0[:8]); put(150, 174, i0[8:])
        put(180, 196, _sbits(q["C_rc"], 16))
        om = _sbits(q["omega"], 32)
        put(196, 204, om[:8]); put(210, 234, om[8:])
        put(240, 264, _sbits(q["omegaDot"], 24))
        put(270, 278, _ubits(q["IODE"], 8))
        put(278, 292, _sbits(q["iDot"], 14))
    return [bits[30 * w:30 * w + 24] for w in range(10)]


def encode_stream(q, first_sf_id, first_tow_s, n_subframes):
    """0/1 bits of ``n_subframes`` consecutive subframes; the first one starts at GPS time
    ``first_tow_s`` (multiple of 6) and has ID ``first_sf_id``."""
    out = []
    d29s, d30s = 0, 0
    for k in range(n_subframes):
        sf_id = (first_sf_id - 1 + k) % 5 + 1
        tow_count = ((first_tow_s + 6 * k) // 6 + 1) % 100800      # HOW carries the start of the NEXT subframe
        for w in _subframe_data(q, sf_id, tow_count):
            word = lnav_word(w, d29s, d30s)
            out.extend(word)
            d29s, d30s = word[28], word[29]
    return np.array(out, dtype=np.int8)


# ------------------------------------------------------------------------------------- geometry
def sat_ecef(e, t):
    """ECEF position (m) at GPS time t from broadcast elements (IS-GPS-200 table 20-IV), no clock terms."""
    a = e["sqrtA"] ** 2
    tk = t - e["t_oe"]
    if tk > 302400:
        tk -= 604800
    if tk < -302400:
        tk += 604800
    n = np.sqrt(GM / a ** 3) + e.get("deltan", 0.0)
    m = e["M_0"] + n * tk
    ecc = e["e"]
    ea = m
    for _ in range(20):
        ea = m + ecc * np.sin(ea)
    nu = np.arctan2(np.sqrt(1 - ecc ** 2) * np.sin(ea), np.cos(ea) - ecc)
    phi = nu + e["omega"]
    u = phi + e.get("C_uc", 0.0) * np.cos(2 * phi) + e.get("C_us", 0.0) * np.sin(2 * phi)
    r = a * (1 - ecc * np.cos(ea)) + e.get("C_rc", 0.0) * np.cos(2 * phi) + e.get("C_rs", 0.0) * np.sin(2 * phi)
    inc = e["i_0"] + e.get("iDot", 0.0) * tk + e.get("C_ic", 0.0) * np.cos(2 * phi) + e.get("C_is", 0.0) * np.sin(2 * phi)
    om = e["omega_0"] + (e.get("omegaDot", 0.0) - OMEGA_E) * tk - OMEGA_E * e["t_oe"]
    xp, yp = r * np.cos(u), r * np.sin(u)
    return np.array([xp * np.cos(om) - yp * np.cos(inc) * np.sin(om),
                     xp * np.sin(om) + yp * np.cos(inc) * np.cos(om),
                     yp * np.sin(inc)])


def llh_to_ecef(lat_deg, lon_deg, h):
    a, f = 6378137.0, 1 / 298.257223563
    e2 = f * (2 - f)
    lat, lon = np.radians(lat_deg), np.radians(lon_deg)
    nn = a / np.sqrt(1 - e2 * np.sin(lat) ** 2)
    return np.array([(nn + h) * np.cos(lat) * np.cos(lon), (nn + h) * np.cos(lat) * np.sin(lon),
                     (nn * (1 - e2) + h) * np.sin(lat)])


def geometric_range(e, t_tx, rx):
    """Range with the Earth rotating during the flight (the model of leastSquarePos / e_r_corr)."""
    x = sat_ecef(e, t_tx)
    rho = np.linalg.norm(x - rx)
    for _ in range(4):
        th = OMEGA_E * rho / C
        xr = np.array([np.cos(th) * x[0] + np.sin(th) * x[1], -np.sin(th) * x[0] + np.cos(th) * x[1], x[2]])
        rho = np.linalg.norm(xr - rx)
    return rho, xr


def elevation(rx, xs):
    up = rx / np.linalg.norm(rx)
    d = xs - rx
    return np.degrees(np.arcsin(np.dot(d, up) / np.linalg.norm(d)))


def build_scenario(seed=2, n_sats=8, fs=38.192e6, f_if=9.548e6, cn0=45.0, rx_llh=(40.0, -105.0, 1600.0),
                   tow=388800, first_boundary_ms=5200.0, min_elev=15.0, sigma=12.0):
    """A receiver at ``rx_llh`` and ``n_sats`` satellites above ``min_elev`` degrees.  The subframe that
    starts at GPS time ``tow`` reaches the antenna ``first_boundary_ms`` (+ range/c differences) after
    sample 0.  Returns (synth.RecordingSpec, truth dict)."""
    rng = np.random.default_rng(seed)
    rx = llh_to_ecef(*rx_llh)
    n_code = int(round(fs / 1000.0))
    sats, ephs = [], []
    prns = list(rng.permutation(np.arange(1, 33)))
    tries = 0
    while len(sats) < n_sats:
        tries += 1
        assert tries < 5000, "could not place the constellation"
        e = dict(sqrtA=5153.6 + rng.uniform(-1, 1), e=rng.uniform(0.001, 0.008), i_0=np.radians(55.0 + rng.uniform(-1, 1)),
                 omega_0=rng.uniform(-np.pi, np.pi), omega=rng.uniform(-np.pi, np.pi), M_0=rng.uniform(-np.pi, np.pi),
                 omegaDot=-8.0e-9, iDot=0.0, deltan=4.5e-9, t_oe=float(tow), weekNumber=2100, IODE=int(rng.integers(1, 255)))
        q = quantize_ephemeris(e)
        eq = dequantize_ephemeris(q)
        rho, xs = geometric_range(eq, float(tow), rx)
        if elevation(rx, xs) < min_elev:
            continue
        rho1, _ = geometric_range(eq, float(tow) + 1.0, rx)
        ephs.append((q, eq, rho, rho1 - rho))
        sats.append(len(sats))
    rho_min = min(x[2] for x in ephs)
    specs = []
    truth = dict(rx=rx, rx_llh=rx_llh, tow=tow, prn=[], range=[], doppler=[], boundary_sample=[], eph=[])
    for k, (q, eq, rho, rdot) in enumerate(ephs):
        prn = int(prns[k])
        doppler = -rdot / (C / synth.L1_HZ)
        n_s = int(round(first_boundary_ms * n_code + (rho - rho_min) / C * fs))
        # stream: the subframe that starts at `tow` gets ID 1 + (k % 5); one more subframe precedes it
        boundary_id = 1 + (k % 5)
        bits01 = encode_stream(q, (boundary_id - 2) % 5 + 1, tow - 6, 8)
        sat = synth.SatSpec(prn, doppler, 0, cn0=cn0, nav_bits=bits01.astype(np.int8) * 2 - 1,
                            carrier_phase=float(rng.uniform(0, 1)))
        sat.frame_sync = (n_s, 300)                  # sample of the boundary, bit index of that boundary
        specs.append(sat)
        truth["prn"].append(prn); truth["range"].append(rho); truth["doppler"].append(doppler)
        truth["boundary_sample"].append(n_s); truth["eph"].append(eq)
    spec = synth.RecordingSpec(specs, fs=fs, f_if=f_if, sigma=sigma, seed=seed, n_bits=2400)
    return spec, truth

"""Golden of BASELINE config 2 (full main.py run: acquisition + 8 channels x 37 000 ms tracking + postNavigation),
made by the REFERENCE's own caller, unmodified except for the mechanical Python-3 shim:

    python tests/golden/make_golden_c2.py            # build container only (needs /root/reference); ~15 min, 1 core

``initialize.Settings.postProcessing`` (initialize.py:454-527) is run on a synthetic LNAV recording
(``navsynth.build_scenario(seed=2)``, 37 100 ms, 1.417 GB, regenerated from the seed -- never stored) with the
reference's own ``acquisition`` / ``tracking`` / ``postNavigation`` modules; the ``.plot()`` methods (matplotlib) are
replaced by stubs that capture the result objects.  Stored, compactly (tests/golden/c2_full.npz, < 1 MB):

  * acquisition: carrFreq, codePhase, peakMetric (32 each) and the channel table of preRun;
  * tracking, every millisecond: absoluteSample as int8 deltas to the nominal 38 192 samples per code period
    (exact), sign(I_P) bit-packed (exact);
  * tracking, every 37th millisecond: all 13 series as float64 (tolerances are written in the test);
  * navigation: the solution arrays of postNavigate (X, Y, Z, dt, DOP, latitude, longitude, height, E, N, U,
    corrected pseudoranges, az/el) for all measurement epochs;
  * SHA-1 of the recording (the device generator must reproduce it byte for byte).
"""
import contextlib
import hashlib
import io
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import make_ref_shim                      # noqa: E402
from softgnss_python_b200 import navsynth, synth      # noqa: E402

N = 38192
MS = 37000
TOTAL_MS = 37100
SEED = 2
SUB = 37
TMP = os.environ.get("SGX_C2_TMP", "/tmp/sgx_c2")
FIELDS = ("absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L",
          "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt")
SOL_FIELDS = ("X", "Y", "Z", "dt", "latitude", "longitude", "height", "E", "N", "U")


def _gen_chunk(args):
    lo, n = args
    spec, _ = navsynth.build_scenario(seed=SEED)
    return synth.generate_cpu(spec, n, start=lo)


def recording_path():
    os.makedirs(TMP, exist_ok=True)
    path = os.path.join(TMP, "rec_c2.bin")
    total = TOTAL_MS * N
    if not (os.path.exists(path) and os.path.getsize(path) == total):
        step = 200 * N
        with mp.get_context("fork").Pool(os.cpu_count()) as pool, open(path, "wb") as f:
            for part in pool.imap(_gen_chunk, [(lo, min(step, total - lo)) for lo in range(0, total, step)]):
                part.tofile(f)
    return path


def file_sha1(path):
    h = hashlib.sha1()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def main():
    t0 = time.time()
    path = recording_path()
    sha = file_sha1(path)
    print("recording %s (%.2f GB, sha1 %s) %.0f s" % (path, os.path.getsize(path) / 1e9, sha, time.time() - t0))
    ref = make_ref_shim.import_ref()
    init, acq_mod, trk_mod, nav_mod = ref["initialize"], ref["acquisition"], ref["tracking"], ref["postNavigation"]
    captured = {}
    acq_mod.AcquisitionResult.plot = lambda self: captured.__setitem__("acq", self)
    trk_mod.TrackingResult.plot = lambda self: captured.__setitem__("trk", self)
    nav_mod.NavigationResult.plot = lambda self: captured.__setitem__("nav", self)
    s = init.Settings()
    s.msToProcess = float(MS)
    s.numberOfChannels = 8
    s.useTropCorr = False          # as tools/e2e_nav.py: the synthetic geometry has no troposphere
    s.fileName = path
    s.plotTracking = False         # postProcessing plots tracking `if not settings.plotTracking` (initialize.py:521)
    cwd = os.getcwd()
    os.chdir(TMP)                  # postProcessing caches trackingResults_python.npy in the working directory
    if os.path.exists("trackingResults_python.npy"):
        os.remove("trackingResults_python.npy")
    t0 = time.time()
    sink = io.StringIO()
    try:
        with contextlib.redirect_stdout(sink):
            s.postProcessing()
    finally:
        os.chdir(cwd)
    print("reference postProcessing: %.0f s" % (time.time() - t0))
    print(sink.getvalue()[-1500:])
    a, t, nav = captured["acq"], captured["trk"], captured["nav"]
    r = t.results
    track = np.stack([np.stack([np.asarray(r[i][f], dtype=np.float64) for f in FIELDS]) for i in range(len(r))])
    abs_s = track[:, 0]
    start = s.skipNumberOfBytes + np.asarray(a.channels.codePhase[:len(r)], dtype=np.float64)
    prev = np.concatenate([start[:, None], abs_s[:, :-1]], axis=1)
    delta = abs_s - prev - N
    assert np.abs(delta).max() < 100 and np.array_equal(delta, np.round(delta))
    sol = nav._solutions[0]
    out = dict(recording_sha1=sha, total_ms=TOTAL_MS, ms=MS, seed=SEED, sub=SUB,
               carrFreq=a.carrFreq, codePhase=a.codePhase, peakMetric=a.peakMetric,
               ch_PRN=np.asarray(a.channels.PRN), ch_acquiredFreq=np.asarray(a.channels.acquiredFreq),
               ch_codePhase=np.asarray(a.channels.codePhase),
               trk_PRN=np.asarray(r.PRN), abs_delta=delta.astype(np.int8),
               ip_sign=np.packbits(track[:, 3] > 0, axis=1), ip_zero=np.argwhere(track[:, 3] == 0),
               sub_series=track[:, :, ::SUB].copy())
    for f in SOL_FIELDS:
        out["sol_" + f] = np.asarray(getattr(sol, f), dtype=np.float64)
    out["sol_DOP"] = np.asarray(sol.DOP, dtype=np.float64)
    ch = sol.channel[0]
    for f in ("rawP", "correctedP", "az", "el", "PRN"):
        out["solch_" + f] = np.asarray(getattr(ch, f), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "c2_full.npz"), **out)
    print("wrote c2_full.npz: %.0f KB; fixes %d" % (os.path.getsize(os.path.join(HERE, "c2_full.npz")) / 1e3,
                                                    int(np.sum(~np.isnan(out["sol_X"])))))


if __name__ == "__main__":
    main()

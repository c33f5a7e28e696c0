"""End-to-end check of BASELINE config 2 (full main.py run on a synthetic 8-satellite file):

    python tools/e2e_nav.py oracle   # build container: CPU generator + oracle acquisition/tracking (parallel)
    python tools/e2e_nav.py gpu      # B200 box: device generator + sgx_acquire/sgx_track  -> gpurun_out/e2e_gpu.npz
    python tools/e2e_nav.py compare  # build container: the reference's own postNavigation (through the shim)
                                     # on both tracking results, diff + distance to the true antenna position

The downstream consumer (postNavigation.py:75-305) is the unmodified reference; it reads I_P,
absoluteSample, PRN and status of the tracking recarray."""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import navsynth, synth                     # noqa: E402
from softgnss_python_b200.settings import Settings                   # noqa: E402

N = 38192
MS = 37000
SEED = 2
TMP = os.environ.get("SGX_E2E_TMP", "/tmp/sgx_e2e")
FIELDS = ("absoluteSample", "codeFreq", "carrFreq", "I_P", "I_E", "I_L", "Q_E", "Q_P", "Q_L",
          "dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt")


def scenario():
    return navsynth.build_scenario(seed=SEED)


def settings():
    s = Settings(msToProcess=float(MS), numberOfChannels=8)
    s.useTropCorr = False
    return s


def _gen_chunk(args):
    lo, n = args
    spec, _ = scenario()
    return synth.generate_cpu(spec, n, start=lo)


def _track_one(args):
    ch = args
    from oracle import gnss_oracle as orc
    data = np.load(os.path.join(TMP, "rec.npy"), mmap_mode="r")
    c = np.load(os.path.join(TMP, "channels.npz"))
    series, done = orc.track_channel(data, c["PRN"][ch], c["acquiredFreq"][ch], c["codePhase"][ch], settings(), MS)
    assert done == MS
    return np.stack([series[f] for f in FIELDS])


def run_oracle():
    from oracle import gnss_oracle as orc
    os.makedirs(TMP, exist_ok=True)
    total = (MS + 100) * N
    t = time.time()
    step = 200 * N
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        parts = pool.map(_gen_chunk, [(lo, min(step, total - lo)) for lo in range(0, total, step)])
    data = np.concatenate(parts)
    np.save(os.path.join(TMP, "rec.npy"), data)
    print("generated %.2f GB in %.0f s" % (data.size / 1e9, time.time() - t))
    s = settings()
    t = time.time()
    acq = orc.acquire(data[:11 * N], s)
    ch = orc.pre_run(acq, s)
    print("oracle acquisition %.0f s: PRN" % (time.time() - t), ch["PRN"].tolist())
    np.savez(os.path.join(TMP, "channels.npz"), PRN=ch["PRN"], acquiredFreq=ch["acquiredFreq"], codePhase=ch["codePhase"])
    t = time.time()
    with mp.get_context("fork").Pool(8) as pool:
        out = pool.map(_track_one, list(range(8)))
    print("oracle tracking 8 x %d ms in %.0f s" % (MS, time.time() - t))
    np.savez_compressed(os.path.join(TMP, "e2e_oracle.npz"), track=np.stack(out), PRN=ch["PRN"],
                        acquiredFreq=ch["acquiredFreq"], codePhase=ch["codePhase"], carrFreq=acq["carrFreq"],
                        peakMetric=acq["peakMetric"])


def run_gpu():
    import torch
    from softgnss_python_b200 import _native
    from softgnss_python_b200.acquisition import AcquisitionResult
    from softgnss_python_b200.tracking import TrackingResult
    spec, truth = scenario()
    L = _native.lib()
    total = (MS + 100) * N
    stride = (total + 15) // 16 * 16
    dev = torch.empty((1, stride), dtype=torch.int8, device="cuda")
    sp, bits = _native.make_synth_specs([spec])
    L.synth(dev, stride, total, 0, sp, bits, synth.cos_lut(), _native.ca_chips_int8(), 0)
    torch.cuda.synchronize()
    s = settings()
    t = time.time()
    a = AcquisitionResult(s)
    a.acquire(dev[0, :11 * N])
    a.preRun()
    a.showChannelStatus()
    tr = TrackingResult(a)
    tr.track(dev[0, :total])
    torch.cuda.synchronize()
    print("GPU acquisition + tracking (8 x %d ms): %.2f s" % (MS, time.time() - t))
    r = tr.results
    track = np.stack([np.stack([np.asarray(r[i][f], dtype=np.float64) for f in FIELDS]) for i in range(len(r))])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "e2e_gpu.npz"), track=track.astype(np.float64),
                        PRN=a.channels.PRN, acquiredFreq=a.channels.acquiredFreq, codePhase=a.channels.codePhase,
                        carrFreq=a.carrFreq, peakMetric=a.peakMetric)


def _nav(track, prn, label):
    """Run the reference's postNavigation (shim) on a tracking result array [ch, 13, ms]."""
    from oracle import make_ref_shim
    ref = make_ref_shim.import_ref()
    s = ref["initialize"].Settings()
    s.msToProcess = float(MS)
    s.numberOfChannels = 8
    s.useTropCorr = False
    dtype = [('status', 'S1')] + [(f, 'object') for f in FIELDS] + [('PRN', 'int64')]
    rec = [(b'T',) + tuple(track[c, i] for i in range(13)) + (int(prn[c]),) for c in range(track.shape[0])]
    res = np.rec.fromrecords(rec, dtype=dtype)

    class T(object):
        results = res
        channels = None
        settings = s
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        nav = ref["postNavigation"].NavigationResult(T())
        nav.postNavigate()
    sol = nav._solutions[0]
    return sol


def compare():
    spec, truth = scenario()
    o = np.load(os.path.join(TMP, "e2e_oracle.npz"))
    g = np.load(os.path.join(ROOT, "gpurun_out", "e2e_gpu.npz"))
    lines = []
    lines.append("PRN order oracle %s gpu %s" % (o["PRN"].tolist(), g["PRN"].tolist()))
    lines.append("acquiredFreq max diff %.3g Hz, codePhase identical %s" %
                 (np.abs(o["acquiredFreq"] - g["acquiredFreq"]).max(), np.array_equal(o["codePhase"], g["codePhase"])))
    to, tg = o["track"], g["track"]
    lines.append("absoluteSample identical: %s ; sign(I_P) identical: %s" %
                 (np.array_equal(to[:, 0], tg[:, 0]), np.array_equal(np.sign(to[:, 3]), np.sign(tg[:, 3]))))
    scale = np.abs(to[:, 3:9]).max()
    dev = np.abs(to[:, 3:9] - tg[:, 3:9]).max(axis=1)
    lines.append("I/Q deviation: median %.2e, 99.9th pct %.2e, max %.2e of full scale %.3g; ms with a chip reassignment (>1e-4): %d of %d" %
                 (np.median(dev) / scale, np.quantile(dev, 0.999) / scale, dev.max() / scale, scale, int((dev > 1e-4 * scale).sum()), dev.size))
    lines.append("carrFreq max diff %.3g Hz, codeFreq max diff %.3g Hz" %
                 (np.abs(to[:, 2] - tg[:, 2]).max(), np.abs(to[:, 1] - tg[:, 1]).max()))
    so = _nav(to, o["PRN"], "oracle")
    sg = _nav(tg, g["PRN"], "gpu")
    rx = truth["rx"]
    n = int(np.sum(~np.isnan(so.X)))
    for name, s_ in (("oracle tracking", so), ("B200 tracking", sg)):
        err = np.sqrt((s_.X[:n] - rx[0]) ** 2 + (s_.Y[:n] - rx[1]) ** 2 + (s_.Z[:n] - rx[2]) ** 2)
        lines.append("%s -> reference postNavigation: %d fixes, 3-D error to true antenna: first %.1f m, mean %.1f m, max %.1f m"
                     % (name, n, err[0], err.mean(), err.max()))
    d = np.sqrt((so.X[:n] - sg.X[:n]) ** 2 + (so.Y[:n] - sg.Y[:n]) ** 2 + (so.Z[:n] - sg.Z[:n]) ** 2)
    lines.append("position difference oracle-path vs B200-path: max %.3g m; raw pseudorange max diff %.3g m" %
                 (d.max(), np.nanmax(np.abs(so.channel[0].rawP - sg.channel[0].rawP))))
    print("\n".join(lines))
    return lines


if __name__ == "__main__":
    {"oracle": run_oracle, "gpu": run_gpu, "compare": compare}[sys.argv[1]]()

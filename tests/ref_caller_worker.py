"""Worker of tests/test_gpu_reference_caller.py (own process: the module names `acquisition` / `tracking` must resolve to
the drop-ins, not to the reference).

The reference's own caller, `initialize.Settings.postProcessing` (initialize.py:454-527), is run unmodified (Python-3
shim only) with `dropin/` first on sys.path: it imports `acquisition` and `tracking` -> the B200 drop-ins, and
`postNavigation` -> the reference's own module, which consumes the drop-ins' recarrays.  The navigation solutions are
written to the path given on the command line."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 38192


def main():
    out_path, work = sys.argv[1], sys.argv[2]
    sys.path.insert(0, ROOT)
    from oracle import make_ref_shim
    shim = make_ref_shim.build()
    import torch
    from softgnss_python_b200 import _native, navsynth, synth
    g = np.load(os.path.join(ROOT, "tests", "golden", "c2_full.npz"), allow_pickle=False)
    ms, total = int(g["ms"]), int(g["total_ms"]) * N
    # the recording file (1.4 GB), generated on the device
    spec, _ = navsynth.build_scenario(seed=int(g["seed"]))
    L = _native.lib()
    stride = (total + 15) // 16 * 16
    dev = torch.empty((1, stride), dtype=torch.int8, device="cuda")
    sp, bits = _native.make_synth_specs([spec])
    L.synth(dev, stride, total, 0, sp, bits, synth.cos_lut(), _native.ca_chips_int8(), 0)
    path = os.path.join(work, "recording.bin")
    with open(path, "wb") as f:
        for lo in range(0, total, 1 << 26):
            f.write(dev[0, lo:min(total, lo + (1 << 26))].cpu().numpy().tobytes())
    del dev
    torch.cuda.empty_cache()
    # drop-ins first, then the shimmed reference (initialize, postNavigation, ephemeris, geoFunctions)
    sys.path.insert(0, shim)
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    import acquisition
    import tracking
    import initialize
    import postNavigation
    assert acquisition.__file__.startswith(os.path.join(ROOT, "dropin")), acquisition.__file__
    assert tracking.__file__.startswith(os.path.join(ROOT, "dropin")), tracking.__file__
    assert postNavigation.__file__.startswith(shim), postNavigation.__file__
    captured = {}
    acquisition.AcquisitionResult.plot = lambda self: captured.__setitem__("acq", self)
    tracking.TrackingResult.plot = lambda self: captured.__setitem__("trk", self)
    postNavigation.NavigationResult.plot = lambda self: captured.__setitem__("nav", self)
    s = initialize.Settings()
    s.msToProcess = float(ms)
    s.numberOfChannels = 8
    s.useTropCorr = False
    s.fileName = path
    s.plotTracking = False
    os.chdir(work)                      # postProcessing caches trackingResults_python.npy in the working directory
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        s.postProcessing()
    text = sink.getvalue()
    assert "Post processing of the signal is over." in text, text[-2000:]
    nav, trk, acq = captured["nav"], captured["trk"], captured["acq"]
    sol = nav._solutions[0]
    ch = sol.channel[0]
    r = trk.results
    np.savez(out_path, X=sol.X, Y=sol.Y, Z=sol.Z, dt=sol.dt, latitude=sol.latitude, longitude=sol.longitude,
             height=sol.height, E=sol.E, N=sol.N, U=sol.U, DOP=sol.DOP, rawP=ch.rawP, correctedP=ch.correctedP,
             carrFreq=acq.carrFreq, codePhase=acq.codePhase, ch_PRN=np.asarray(acq.channels.PRN),
             absoluteSample=np.stack([np.asarray(x, dtype=np.float64) for x in r.absoluteSample]),
             stdout_tail=text[-600:])


if __name__ == "__main__":
    main()

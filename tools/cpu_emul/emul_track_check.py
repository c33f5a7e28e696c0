"""Developer check: run the tracking / synth kernels under the CPU fiber emulator and diff against
the golden vectors of the reference.  Not part of the product or of the default test run."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native, synth            # noqa: E402
from softgnss_python_b200.settings import to_pod            # noqa: E402
from tests.cases import CASES, N, build_recording, case_settings   # noqa: E402

L = _native.Lib(os.environ.get("SGX_EMUL_LIB", os.path.join(ROOT, "tools", "cpu_emul", "libsoftgnss_emul.so")))


def check_synth():
    case = CASES["trk_skip"]
    spec, data = build_recording(case)
    n = 3 * N + 5
    specs, bits = _native.make_synth_specs([spec])
    out = np.zeros((1, n + 11), dtype=np.int8)
    L.synth(out, out.strides[0], n, 1234, specs, bits, synth.cos_lut(), _native.ca_chips_int8())
    ref = synth.generate_cpu(spec, n, start=1234)
    print("synth identical:", np.array_equal(out[0, :n], ref), "tail untouched:", not out[0, n:].any())


def check_track(name, ms=None):
    case = CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    spec, data = build_recording(case)
    s = case_settings(case)
    if ms:
        s.msToProcess = ms
    pod = to_pod(s)
    chans = _native.make_channels(g["ch_PRN"], g["ch_acquiredFreq"], g["ch_codePhase"])
    c = len(g["ch_PRN"])
    out = np.zeros((1, c, 13, pod.msToProcess))
    t = time.time()
    rc, done = L.track(data.reshape(1, -1), data.size, [data.size], chans, pod, _native.ca_chips_int8(), out)
    print(name, "rc", rc, "done", done.tolist(), "%.1fs" % (time.time() - t))
    m = pod.msToProcess
    act = [i for i in range(c) if g["ch_PRN"][i] != 0]
    for fi, f in enumerate(_native.TRACK_FIELDS):
        ref = g["trk_" + f][:, :m]
        got = out[0, act, fi, :]
        d = np.abs(got - ref)
        rel = d.max() / max(np.abs(ref).max(), 1e-300)
        print("  %-14s max|d|=%.3e rel=%.2e" % (f, d.max(), rel))
    sgn = np.sign(out[0, act, 3]) == np.sign(g["trk_I_P"][:, :m])
    print("  sign(I_P) identical:", bool(sgn.all()), " absoluteSample identical:",
          np.array_equal(out[0, act, 0], g["trk_absoluteSample"][:, :m]))


if __name__ == "__main__":
    check_synth()
    check_track("trk_small", ms=int(sys.argv[1]) if len(sys.argv) > 1 else 40)


def dump(name, ms):
    case = CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    spec, data = build_recording(case)
    s = case_settings(case)
    s.msToProcess = ms
    pod = to_pod(s)
    chans = _native.make_channels(g["ch_PRN"], g["ch_acquiredFreq"], g["ch_codePhase"])
    c = len(g["ch_PRN"])
    out = np.zeros((1, c, 13, pod.msToProcess))
    rc, done = L.track(data.reshape(1, -1), data.size, [data.size], chans, pod, _native.ca_chips_int8(), out)
    for ch in range(c):
        for k in range(ms):
            print(ch, k, "I_P", out[0, ch, 3, k], g["trk_I_P"][ch, k], "Q_P", out[0, ch, 7, k], g["trk_Q_P"][ch, k],
                  "I_E", out[0, ch, 4, k], g["trk_I_E"][ch, k])


def worst(name, ms):
    case = CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    spec, data = build_recording(case)
    s = case_settings(case)
    s.msToProcess = ms
    pod = to_pod(s)
    chans = _native.make_channels(g["ch_PRN"], g["ch_acquiredFreq"], g["ch_codePhase"])
    c = len(g["ch_PRN"])
    out = np.zeros((1, c, 13, pod.msToProcess))
    rc, done = L.track(data.reshape(1, -1), data.size, [data.size], chans, pod, _native.ca_chips_int8(), out)
    d = np.abs(out[0, :, 3, :] - g["trk_I_P"][:, :ms])
    for ch in range(c):
        k = np.argsort(-d[ch])[:5]
        print(ch, [(int(x), float(d[ch, x])) for x in k])

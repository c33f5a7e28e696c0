"""Scratch timing of the tracking kernel on device-generated recordings (developer tool)."""
import sys, time, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native, synth
from softgnss_python_b200.settings import Settings, to_pod

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
MS = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
N = 38192
L = _native.lib()
specs = [synth.RecordingSpec(synth.default_constellation(2000 + r, 8), seed=2000 + r) for r in range(R)]
n = (MS + 2) * N
stride = (n + 15) // 16 * 16
dev = torch.empty((R, stride), dtype=torch.int8, device="cuda")
sp, bits = _native.make_synth_specs(specs)
stream = torch.cuda.current_stream().cuda_stream
t = time.time(); L.synth(dev, stride, n, 0, sp, bits, synth.cos_lut(), _native.ca_chips_int8(), stream); torch.cuda.synchronize()
print("synth %.1f GB in %.3f s" % (R * n / 1e9, time.time() - t))
s = Settings(msToProcess=float(MS))
pod = to_pod(s)
prn, freq, cph = [], [], []
for spc in specs:
    for i, x in enumerate(spc.sats):
        prn.append(x.prn); freq.append(spc.true_carr_freq(i) - 30.0); cph.append((x.code_phase + 1) % N)
ch = _native.make_channels(prn, freq, cph)
out = torch.zeros((R, 8, 13, MS), dtype=torch.float64, device="cuda")
for stage in ("bulk", "cpasync"):
    os.environ["SGX_TRK_STAGE"] = stage
    for it in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc, done = L.track(dev, stride, [n] * R, ch, pod, _native.ca_chips_int8(), out, stream)
        e1.record(); torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) / 1e3
        print(stage, "rc", rc, "done", int(done.min()), "%.4f s -> %.3e channel-ms/s, %.2f us/ms/channel-step, algo %.1f GB/s"
              % (dt, R * 8 * MS / dt, dt / MS * 1e6, R * 8 * MS * 38296 / dt / 1e9))
o = out.cpu().numpy()
ip = o[:, :, 3, :]
print("lock check: mean |I_P| / mean |Q_P| =", np.abs(ip[:, :, 200:]).mean() / np.abs(o[:, :, 7, 200:]).mean())

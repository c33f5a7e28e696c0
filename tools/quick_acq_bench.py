"""Scratch timing of the acquisition pipeline (developer tool)."""
import sys, time, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native, synth
from softgnss_python_b200.settings import Settings
from softgnss_python_b200.acquisition import acquire_batch
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = 38192
L = _native.lib()
specs = [synth.RecordingSpec(synth.default_constellation(1000 + r, 8), seed=1000 + r) for r in range(R)]
n = 11 * N
dev = torch.empty((R, n), dtype=torch.int8, device="cuda")
sp, bits = _native.make_synth_specs(specs)
stream = torch.cuda.current_stream().cuda_stream
L.synth(dev, n, n, 0, sp, bits, synth.cos_lut(), _native.ca_chips_int8(), stream)
s = Settings()
for it in range(4):
    l0 = L.launches()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    r = acquire_batch(dev, s, stream=stream)
    e1.record(); torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / 1e3
    cells = R * 32 * 29 * N
    print("R=%d %.4f s -> %.3e cells/s, detected %d, launches %d" % (R, dt, cells / dt, int((r["carrFreq"] > 0).sum()), L.launches() - l0))

"""BASELINE config 5: acquisition throughput over the sampling rate (4 ... 64 MHz real int8) and the coherent
integration length (1, 2, 5, 10 ms) on one GPU.  Writes one JSON document (default profiles/sweep_r2.json).

    python tools/sweep_c5.py [out.json] [recordings]

Every point: `recordings` synthetic recordings (8 satellites at 45 dB-Hz) generated on the device, reference settings
otherwise (2 blocks with pick-max, 29 Doppler bins of 500 Hz, 32 PRNs, fine search on 10 ms); cells = recordings x 32 x
29 x samplesPerCode (one code period of code phases is searched whatever the coherent length).  Timing: CUDA events,
3 warm-up + 5 timed batches.  `engine` says which transform path ran: the prime-factor kernel exists for
N = 38 192 (31x7x16x11) and N = 16 368 (31x3x16x11) at 1 ms; every other length goes through the generic mixed-radix
engine (run-time radices)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softgnss_python_b200 import _native, synth                     # noqa: E402
from softgnss_python_b200.acquisition import acquire_batch          # noqa: E402
from softgnss_python_b200.settings import Settings                  # noqa: E402

FS = (4.0e6, 8.0e6, 16.0e6, 16.3676e6, 32.0e6, 38.192e6, 64.0e6)
COH = (1, 2, 5, 10)


def factor(n):
    f, p = [], 2
    while n > 1:
        while n % p == 0:
            f.append(p); n //= p
        p += 1
    return f


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sweep_r2.json")
    recs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    L = _native.lib()
    stream = torch.cuda.current_stream().cuda_stream
    chips, lut = _native.ca_chips_int8(), synth.cos_lut()
    points = []
    for fs in FS:
        for coh in COH:
            s = Settings(samplingFreq=fs, IF=fs / 4.0, acqCoherentMs=coh)
            n1 = s.samplesPerCode
            n_ms = max(11, 2 * coh)
            ns = n_ms * n1
            specs = [synth.RecordingSpec(synth.default_constellation(3000 + r, 8, fs=fs, n_code=n1), fs=fs, f_if=fs / 4.0,
                                         seed=3000 + r) for r in range(recs)]
            dev = torch.empty((recs, ns), dtype=torch.int8, device="cuda")
            sp, bits = _native.make_synth_specs(specs)
            L.synth(dev, ns, ns, 0, sp, bits, lut, chips, stream)
            try:
                for _ in range(3):
                    res = acquire_batch(dev, s, stream=stream)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    res = acquire_batch(dev, s, stream=stream)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                cells = recs * 32 * 29 * n1
                pt = dict(fs_mhz=fs / 1e6, samples_per_code=n1, coherent_ms=coh, transform_length=n1 * coh,
                          factors=factor(n1 * coh), recordings=recs, ms_per_batch=ms, cells_per_s=cells / (ms / 1e3),
                          detected=int((res["carrFreq"] > 0).sum()), present=8 * recs,
                          engine="prime-factor kernel (31x7x16x11)" if (n1 * coh == 38192) else
                          "prime-factor kernel (31x3x16x11)" if (n1 * coh == 16368) else "generic mixed-radix passes")
            except _native.NativeError as e:
                pt = dict(fs_mhz=fs / 1e6, samples_per_code=n1, coherent_ms=coh, error=str(e))
            points.append(pt)
            print(json.dumps(pt), flush=True)
            del dev
    doc = dict(what="BASELINE config 5 sweep, one B200, acquisition search cells/s", points=points,
               gpu=torch.cuda.get_device_name(0))
    with open(out_path, "w") as f:
        json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()

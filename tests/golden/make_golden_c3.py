"""Golden of BASELINE config 3 (weak-signal acquisition: 10 ms coherent / 10 x 1 ms blocks, 100 Hz Doppler step = 141
bins, 32 PRNs, 30-35 dB-Hz):

    python tests/golden/make_golden_c3.py          # build container; ~10 min on 4 cores

The reference hard-codes two 1 ms blocks and a 500 Hz grid (acquisition.py:55-57, :68, :101, :129-133), so this
configuration has no reference behaviour; it is pinned by the oracle's parametrised restatement
(oracle/gnss_oracle.py:acquire), which IS the reference at (1 ms, 2 blocks, 500 Hz) -- tests/test_oracle_golden.py.
The oracle needs minutes per recording at these sizes (4 512 complex128 transforms of 381 920 points), hence the
committed fixture: tests/golden/acq_c3.npz holds carrFreq / codePhase / peakMetric of the oracle for two seeded
recordings x two modes; the recordings are regenerated from the seed (SHA-1 stored).

Satellites: nine per recording at 30 ... 43 dB-Hz, so that both modes have detections, misses and PRNs whose
peakMetric lies within 10 % of acqThreshold = 2.5 (a float32-vs-float64 decision flip would show there).
"""
import hashlib
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from softgnss_python_b200 import synth                     # noqa: E402
from softgnss_python_b200.settings import Settings          # noqa: E402

N = 38192
MODES = {"coh10": dict(acqCoherentMs=10, acqNonCoherentBlocks=1, acqDopplerStep=100.0),
         "blk10": dict(acqCoherentMs=1, acqNonCoherentBlocks=10, acqDopplerStep=100.0)}
SEEDS = (1000, 1001)
CN0 = (30.0, 31.0, 32.0, 33.0, 34.0, 35.0, 38.0, 41.0, 43.0)


def recording_spec(seed):
    rng = np.random.default_rng(seed)
    prns = rng.permutation(np.arange(1, 33))[:len(CN0)]
    sats = []
    for p, cn0 in zip(prns, CN0):
        cp = int(rng.integers(40, N - 40))
        dop = float(np.round(rng.uniform(-6500.0, 6500.0), 1))
        sats.append(synth.SatSpec(int(p), dop, cp, cn0=cn0, bit_offset_ms=int(rng.integers(0, 20)),
                                  carrier_phase=float(rng.uniform())))
    return synth.RecordingSpec(sats, seed=seed)


def settings_for(mode):
    return Settings(**MODES[mode])


def _run(args):
    seed, mode = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import gnss_oracle as orc
    spec = recording_spec(seed)
    data = synth.generate_cpu(spec, 11 * N)
    s = settings_for(mode)
    m = MODES[mode]
    r = orc.acquire(data, s, coherent_ms=m["acqCoherentMs"], noncoh_blocks=m["acqNonCoherentBlocks"],
                    doppler_step=m["acqDopplerStep"], clamp_window=True)
    return seed, mode, hashlib.sha1(data.tobytes()).hexdigest(), r


def main():
    jobs = [(seed, mode) for seed in SEEDS for mode in MODES]
    with mp.get_context("fork").Pool(4) as pool:
        res = pool.map(_run, jobs)
    out = {}
    for seed, mode, sha, r in res:
        k = "%s_%d_" % (mode, seed)
        out[k + "sha1"] = sha
        for f in ("carrFreq", "codePhase", "peakMetric"):
            out[k + f] = r[f]
        spec = recording_spec(seed)
        present = sorted(int(p) for p in spec.prn)
        det = (np.nonzero(r["carrFreq"])[0] + 1).tolist()
        near = [(i + 1, round(float(v), 3)) for i, v in enumerate(r["peakMetric"]) if 2.25 <= v <= 2.75]
        print(mode, seed, "present", present, "detected", det, "within 10% of the threshold", near)
    np.savez_compressed(os.path.join(HERE, "acq_c3.npz"), **out)


if __name__ == "__main__":
    main()

"""On-disk -> HBM ingest (replaces fid.seek / np.fromfile of tracking.py:107, :154 and the
dataType / skipNumberOfBytes handling of initialize.py:102, :466-481): `tracking(fid, ...)` on a file streams it with
pread() through two pinned staging buffers into HBM while the kernel tracks what is already resident.

  * a recording of more than 1 GB on disk, non-zero skipNumberOfBytes: results bit-identical to the HBM-resident run,
    peak resident memory of the process grows by less than two staging buffers + the result (never by the file size);
  * only the window the channels can touch is read;
  * dataType int16: same results as the int8 file; a sample outside the int8 range is rejected;
  * an `e2e_file` figure (channel-ms/s from a page-cached file) is written to gpurun_out/ when that directory exists.
"""
import json
import os
import resource
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 38192
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _channels(spec, n_ch):
    prn = np.array([x.prn for x in spec.sats[:n_ch]], dtype=np.int64)
    freq = np.array([spec.true_carr_freq(i) - 30.0 for i in range(n_ch)])
    cph = np.array([(x.code_phase + 1) % N for x in spec.sats[:n_ch]], dtype=np.float64)
    return np.rec.fromarrays([prn, freq, cph, ['T'] * n_ch], names="PRN,acquiredFreq,codePhase,status")


def _series(res):
    from softgnss_python_b200._native import TRACK_FIELDS
    return np.stack([np.stack([np.asarray(res[c][f], dtype=np.float64) for f in TRACK_FIELDS]) for c in range(len(res))])


def test_large_file_is_streamed_and_bit_identical(tmp_path):
    import torch
    from softgnss_python_b200 import _native, synth, tracking as trk
    from softgnss_python_b200.settings import Settings
    ms, skip, n_ch = 27800, 5008, 4
    total = (ms + 400) * N                                  # 1.08 GB of samples; the tail is never needed
    spec = synth.RecordingSpec(synth.default_constellation(77, n_ch, cn0=47.0), seed=77)
    L = _native.lib()
    stride = (total + 15) // 16 * 16
    dev = torch.empty((1, stride), dtype=torch.int8, device="cuda")
    sp, bits = _native.make_synth_specs([spec])
    L.synth(dev, stride, total, 0, sp, bits, synth.cos_lut(), _native.ca_chips_int8(), 0)
    path = str(tmp_path / "recording.bin")
    with open(path, "wb") as f:
        f.write(bytes(skip))                                # skipNumberOfBytes worth of other data in front
        for lo in range(0, total, 1 << 26):
            f.write(dev[0, lo:min(total, lo + (1 << 26))].cpu().numpy().tobytes())
    assert os.path.getsize(path) > (1 << 30)
    s = Settings(msToProcess=float(ms), numberOfChannels=n_ch, skipNumberOfBytes=skip)
    ch = _channels(spec, n_ch)
    # resident run: the same bytes in HBM (with the prefix, so that positions agree)
    whole = torch.zeros(skip + total, dtype=torch.int8, device="cuda")
    whole[skip:] = dev[0, :total]
    del dev
    ref, _ = trk.tracking(whole, ch, s)
    del whole
    torch.cuda.empty_cache()
    chunk_ms = 256
    trk.FILE_CHUNK_SAMPLES = chunk_ms * N                   # 9.8 MB staging buffers
    try:
        rss0 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss          # KB, high-water mark so far
        t0 = time.perf_counter()
        with open(path, "rb") as fid:
            got, _ = trk.tracking(fid, ch, s)
        dt = time.perf_counter() - t0
        rss1 = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    finally:
        trk.FILE_CHUNK_SAMPLES = 0
    assert got is not None and len(got) == n_ch
    assert np.array_equal(_series(got), _series(ref)), "file ingest differs from the resident run"
    result_kb = n_ch * 13 * ms * 8 * 3 / 1024               # the result, its recarray copy and the comparison copy
    grow_kb = rss1 - rss0
    assert grow_kb < 2 * chunk_ms * N / 1024 + result_kb + 32 * 1024, \
        "resident memory grew by %.0f MB while tracking a %.2f GB file" % (grow_kb / 1024, os.path.getsize(path) / 1e9)
    # only the window is read: first sample = skip + min codePhase (rounded down to 16), msToProcess periods (+ margin)
    pod = __import__("softgnss_python_b200.settings", fromlist=["to_pod"]).to_pod(s)
    chans = _native.make_channels([int(x) for x in ch.PRN], [float(x) for x in ch.acquiredFreq], [float(x) for x in ch.codePhase])
    out = np.zeros((1, n_ch, 13, ms))
    rc, done, win = L.track_file(path, 1, chans, pod, _native.ca_chips_int8(), out)
    assert rc == 0 and (done == ms).all()
    assert win[0] == (skip + int(ch.codePhase.min())) // 16 * 16
    assert win[1] < (ms + 4) * N * 1.001 + int(ch.codePhase.max() - ch.codePhase.min())
    assert np.array_equal(out[0], _series(ref))
    rec = dict(metric="tracking channel-ms/s from a file (page cache, pread -> pinned double buffer -> HBM)",
               value=n_ch * ms / dt, seconds=dt, file_gb=os.path.getsize(path) / 1e9, channels=n_ch, ms=ms,
               read_gbs=win[1] / dt / 1e9, rss_growth_mb=grow_kb / 1024)
    print(json.dumps(rec))
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", "e2e_file.json"), "w") as f:
            json.dump(rec, f)


def test_int16_files_and_rejections(tmp_path):
    from softgnss_python_b200 import _native, synth, tracking as trk
    from softgnss_python_b200.settings import Settings
    ms, n_ch, skip = 40, 3, 1000
    spec = synth.RecordingSpec(synth.default_constellation(78, n_ch, cn0=48.0), seed=78)
    data = synth.generate_cpu(spec, (ms + 3) * N)
    ch = _channels(spec, n_ch)
    p8, p16 = str(tmp_path / "a.int8"), str(tmp_path / "a.int16")
    with open(p8, "wb") as f:
        f.write(bytes(skip)); f.write(data.tobytes())
    with open(p16, "wb") as f:
        f.write(bytes(2 * skip)); f.write(data.astype("<i2").tobytes())
    s8 = Settings(msToProcess=float(ms), numberOfChannels=n_ch, skipNumberOfBytes=skip)
    s16 = Settings(msToProcess=float(ms), numberOfChannels=n_ch, skipNumberOfBytes=2 * skip, dataType="int16")
    r8, _ = trk.tracking(p8, ch, s8)
    with open(p16, "rb") as fid:
        r16, _ = trk.tracking(fid, ch, s16)
    assert np.array_equal(_series(r8), _series(r16))
    mem, _ = trk.tracking(np.concatenate([np.zeros(skip, dtype=np.int8), data]), ch, s8)
    assert np.array_equal(_series(r8), _series(mem))
    bad = data.astype("<i2")
    bad[5 * N] = 300
    with open(p16, "wb") as f:
        f.write(bytes(2 * skip)); f.write(bad.tobytes())
    with pytest.raises(_native.NativeError):
        trk.tracking(p16, ch, s16)
    with pytest.raises(_native.NativeError):
        trk.tracking(p8, ch, Settings(msToProcess=float(ms), numberOfChannels=n_ch, dataType="float32"))
    # a file that ends early: the reference's message and None (tracking.py:159-163)
    with open(p8, "r+b") as f:
        f.truncate(skip + 20 * N)
    res, _ = trk.tracking(p8, ch, s8)
    assert res is None

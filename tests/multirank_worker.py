"""Worker of tests/test_gpu_multirank.py: one process per GPU under torchrun (NCCL).

Every rank computes (a) its shard through the sharded entry points of softgnss_python_b200.dist -- acquisition split
by PRN, tracking and the downstream navigation chain split by recording, results gathered GPU to GPU over NCCL --
and (b) the whole problem alone on its own GPU, and asserts that the two are BYTE-IDENTICAL (SURVEY.md section 8(e):
"G-GPU results byte-identical to 1-GPU results").  Writes one JSON report per rank into $SGX_MR_OUT."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N = 38192


def main():
    import torch
    import torch.distributed as dist
    from softgnss_python_b200 import _native, navsynth, synth
    from softgnss_python_b200 import dist as sd
    from softgnss_python_b200.acquisition import acquire_batch, preRun
    from softgnss_python_b200.settings import Settings
    from softgnss_python_b200.tracking import track_batch
    from softgnss_python_b200 import postnav
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _native.lib()
    L.check(L.dll.sgx_set_device(local))
    stream = torch.cuda.current_stream().cuda_stream
    chips, lut = _native.ca_chips_int8(), synth.cos_lut()
    report = {"rank": rank, "world": world}

    def generate(specs, n):
        stride = (n + 15) // 16 * 16
        dev = torch.empty((len(specs), stride), dtype=torch.int8, device="cuda")
        sp, bits = _native.make_synth_specs(specs)
        L.synth(dev, stride, n, 0, sp, bits, lut, chips, stream)
        return dev, stride

    # ---- acquisition split by PRN: 3 recordings, 32 PRNs over the ranks ------------------------------------------
    specs = [synth.RecordingSpec(synth.default_constellation(500 + r, 8), seed=500 + r) for r in range(3)]
    dev, stride = generate(specs, 11 * N)
    sig = dev[:, :11 * N]
    s = Settings()
    whole = acquire_batch(sig, s, stream=stream)
    t0 = time.perf_counter()
    shard = sd.acquire_sharded(sig, s, stream=stream)
    report["acquire_sharded_s"] = time.perf_counter() - t0
    for f in ("carrFreq", "codePhase", "peakMetric"):
        assert whole[f].shape == (3, 32) and np.array_equal(whole[f], shard[f]), "acquisition " + f
    report["acq_detected"] = int((whole["carrFreq"] > 0).sum())

    # ---- tracking split by recording: 2 x world recordings x 4 channels x 200 ms --------------------------------
    n_rec, ms = 2 * world, 200
    ts = Settings(msToProcess=float(ms), numberOfChannels=4)
    tspecs = [synth.RecordingSpec(synth.default_constellation(700 + r, 4), seed=700 + r) for r in range(n_rec)]
    tdev, tstride = generate(tspecs, (ms + 3) * N)
    acq = acquire_batch(tdev[:, :11 * N], ts, stream=stream)
    chans = []
    for r in range(n_rec):
        rec = np.rec.fromarrays([acq["carrFreq"][r], acq["codePhase"][r], acq["peakMetric"][r]],
                                names="carrFreq,codePhase,peakMetric")
        chans.append(preRun(rec, ts))
    rec_len = [(ms + 3) * N] * n_rec
    out_all = torch.zeros((n_rec, 4, 13, ms), dtype=torch.float64, device="cuda")
    rc, out_all, done_all = track_batch(tdev, rec_len, chans, ts, out=out_all, stream=stream)
    assert rc == 0
    lo, hi = sd.shard_range(n_rec, rank, world)
    rc, out_g, done_g = sd.track_sharded(tdev[lo:hi], rec_len[lo:hi], chans[lo:hi], ts, stream=stream)
    assert rc == 0 and out_g.is_cuda, "the gathered result must stay on the device"
    assert torch.equal(out_g, out_all), "tracking result differs between the sharded and the single-GPU run"
    assert np.array_equal(np.asarray(done_g), np.asarray(done_all))
    report["track_channels"] = int((np.asarray(done_all) == ms).sum())
    # uneven shards (rank 0 takes one recording more) and the keep-local mode
    cuts = [0] + [min(n_rec, 1 + (r + 1) * (n_rec - 1) // world) for r in range(world)]
    cuts[-1] = n_rec
    a, b = cuts[rank], cuts[rank + 1]
    rc, out_u, _ = sd.track_sharded(tdev[a:b], rec_len[a:b], chans[a:b], ts, stream=stream)
    assert rc == 0 and torch.equal(out_u, out_all), "uneven shards"
    rc, out_l, _ = sd.track_sharded(tdev[lo:hi], rec_len[lo:hi], chans[lo:hi], ts, stream=stream, gather=False)
    assert rc == 0 and torch.equal(out_l, out_all[lo:hi]), "gather=False keeps the local shard"
    # a short recording on ONE rank must be seen by every rank
    short = list(rec_len[lo:hi])
    if rank == world - 1:
        short[-1] = 50 * N
    rc, _, _ = sd.track_sharded(tdev[lo:hi], short, chans[lo:hi], ts, stream=stream, gather=False)
    assert rc == _native.SGX_ERR_SHORT, "status of the failing rank must reach all ranks (got %d)" % rc
    del tdev, out_all, out_g, out_u, out_l

    # ---- full chain split by recording: `world` LNAV recordings x 8 channels x 37 000 ms ----------------------------
    ms = 37000
    ns = Settings(msToProcess=float(ms))
    ns.useTropCorr = False
    n = (ms + 2) * N
    scen = [navsynth.build_scenario(seed=2 + r) for r in range(world)]
    ndev, nstride = generate([x[0] for x in scen], n)

    def channel_table(spec):
        prn = [x.prn for x in spec.sats]
        freq = [spec.true_carr_freq(i) - 20.0 for i in range(len(prn))]
        cph = [float((x.code_phase + 1) % N) for x in spec.sats]
        return np.rec.fromarrays([np.array(prn, dtype=np.int64), freq, cph, ['T'] * len(prn)],
                                 names="PRN,acquiredFreq,codePhase,status")
    nch = [channel_table(x[0]) for x in scen]
    prn_all = np.stack([c.PRN for c in nch])
    out_n = torch.zeros((world, 8, 13, ms), dtype=torch.float64, device="cuda")
    rc, out_n, done_n = track_batch(ndev, [n] * world, nch, ns, out=out_n, stream=stream)
    assert rc == 0 and int(np.asarray(done_n).min()) == ms
    nav_all = postnav.post_navigate_batch(out_n, prn_all, ns, stream=stream)
    t0 = time.perf_counter()
    rc, out_r, _ = sd.track_sharded(ndev[rank:rank + 1], [n], nch[rank:rank + 1], ns, stream=stream, gather=False)
    assert rc == 0 and torch.equal(out_r, out_n[rank:rank + 1])
    nav_g = sd.post_navigate_sharded(out_r, prn_all[rank:rank + 1], ns, world, stream=stream)
    report["nav_chain_sharded_s"] = time.perf_counter() - t0
    for k in ("sol", "rawP", "correctedP", "el", "az", "active", "n_epochs", "subFrameStart", "ready", "tow"):
        assert np.array_equal(np.asarray(nav_g[k]), np.asarray(nav_all[k]), equal_nan=True), "navigation " + k
    report["nav_epochs"] = [int(x) for x in nav_all["n_epochs"]]
    assert min(report["nav_epochs"]) > 50
    # timed gather of a tracking result of this size (8 channels x 13 x 37 000 doubles per recording)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g = sd.gather_results(out_r, world, axis=0)
    e1.record()
    torch.cuda.synchronize()
    assert torch.equal(g, out_n)
    report["track_gather_ms"] = e0.elapsed_time(e1)
    report["track_gather_bytes_per_rank"] = int(out_r.numel() * 8)
    report["ok"] = True
    with open(os.path.join(os.environ["SGX_MR_OUT"], "rank%d.json" % rank), "w") as f:
        json.dump(report, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

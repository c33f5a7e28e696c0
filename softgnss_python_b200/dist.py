"""Multi-GPU sharding of the two hot paths (one process per GPU, ``torch.distributed``).

Both paths consist of independent units: acquisition = (recording, PRN),
tracking = (recording, channel).  Units are split into contiguous ranges per rank, every rank runs the
single-GPU C-ABI call on its range, and the only collective is the gather of the results
(``all_gather`` over NCCL/NVLink on the GPU box, gloo in the CPU tests).  No exchange happens inside
either algorithm, so an N-GPU run returns exactly the bytes of the 1-GPU run.
"""
import numpy as np


def shard_range(n_units, rank, world):
    """Contiguous, balanced [lo, hi) of rank's units; the union over ranks is [0, n_units)."""
    base, extra = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist


def gather_results(local, n_units, axis=0):
    """All ranks contribute their shard (numpy array whose ``axis`` has the rank's unit count); every
    rank gets the full array in unit order.  Works with any backend (tensors are moved to the
    backend's device)."""
    import torch
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    local = np.moveaxis(np.ascontiguousarray(local), axis, 0)
    counts = [shard_range(n_units, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in counts)
    pad = np.zeros((width,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    mine = torch.from_numpy(pad).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.concatenate([p.cpu().numpy()[:hi - lo] for p, (lo, hi) in zip(parts, counts)], axis=0)
    return np.moveaxis(out, 0, axis)


def acquire_sharded(signals, settings, stream=0):
    """Acquisition of R recordings split by PRN over the ranks (north star: "acquisition is split by
    PRN"): every rank searches its PRN range on all recordings; results gathered to all ranks."""
    from .acquisition import acquire_batch
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    nsat = min(32, len(settings.acqSatelliteList))
    lo, hi = shard_range(nsat, rank, world)
    res = acquire_batch(signals, settings, prn_first=lo, prn_count=hi - lo, stream=stream) if hi > lo else \
        dict(carrFreq=np.zeros((signals.shape[0], 0)), codePhase=np.zeros((signals.shape[0], 0)),
             peakMetric=np.zeros((signals.shape[0], 0)))
    return {k: gather_results(v, nsat, axis=1) for k, v in res.items()}


def track_sharded(recordings, rec_len, channel_sets, settings, stream=0):
    """Tracking of this rank's recordings (the caller passes only the local shard: recordings live on
    the GPU that tracks them); returns the gathered ``out [R_total, C, 13, ms]`` and ``ms_done``."""
    from .tracking import track_batch
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rc, out, done = track_batch(recordings, rec_len, channel_sets, settings, stream=stream)
    if hasattr(out, "cpu"):
        out = out.cpu().numpy()
    if world == 1:
        return rc, out, done
    import torch
    n_local = torch.tensor([len(channel_sets)])
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local.to("cuda") if dist.get_backend() == "nccl" else n_local)
    total = int(sum(int(c.item()) for c in counts))
    assert all(int(c.item()) == len(channel_sets) for c in counts), "equal shards expected"
    return rc, gather_results(out, total, axis=0), gather_results(done, total, axis=0)


def gather_epochs(local, n_units):
    """Gather per-recording arrays whose second axis is the measurement epoch (``[R_local, E_local, ...]``): the
    epoch count differs between ranks (it follows from each recording's subframe start, postNavigation.py:199), so
    the shards are first padded to the largest E -- NaN for floating point (the reference's initial value of a
    solution column), 0 otherwise -- and then gathered in recording order."""
    import torch
    dist = _dist()
    local = np.ascontiguousarray(local)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    e_max = torch.tensor([local.shape[1]], device=device)
    dist.all_reduce(e_max, op=dist.ReduceOp.MAX)
    e_max = int(e_max.item())
    fill = np.nan if local.dtype.kind == "f" else 0
    pad = np.full((local.shape[0], e_max) + local.shape[2:], fill, dtype=local.dtype)
    pad[:, :local.shape[1]] = local
    return gather_results(pad, n_units, axis=0)


def post_navigate_sharded(track_out, prn, settings, n_total, stream=0):
    """Preamble search -> ephemeris decoding -> measurement loop for this rank's recordings (``track_out`` is the
    local shard of ``track_batch``'s output and stays on its GPU); the solutions of all ``n_total`` recordings are
    gathered to every rank.  No exchange inside the chain: recordings are independent."""
    from . import postnav
    out = postnav.post_navigate_batch(track_out, prn, settings, stream=stream)
    res = {k: gather_epochs(out[k], n_total) for k in ("sol", "rawP", "correctedP", "el", "az", "active")}
    for k in ("n_epochs", "subFrameStart", "ready", "tow"):
        res[k] = gather_results(out[k], n_total, axis=0)
    return res

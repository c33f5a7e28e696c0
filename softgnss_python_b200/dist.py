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


def _counts(n_units, world, counts):
    if counts is None:
        return [shard_range(n_units, r, world) for r in range(world)]
    bounds, lo = [], 0
    for c in counts:
        bounds.append((lo, lo + int(c)))
        lo += int(c)
    assert lo == n_units, "shard sizes do not add up to the number of units"
    return bounds


def gather_results(local, n_units, axis=0, counts=None):
    """All ranks contribute their shard (array whose ``axis`` has the rank's unit count); every rank gets the full
    array in unit order.  ``local`` may be a numpy array (result: numpy) or a torch tensor (result: tensor on the same
    device).  A CUDA tensor under NCCL never leaves the device: the shards travel GPU -> GPU over NVLink straight
    into their slice of the output -- one ``all_gather_into_tensor`` when the shards are equal, one broadcast per
    rank otherwise (no padding to the widest shard, no host staging).  ``counts``: explicit units per rank
    (default: ``shard_range``)."""
    import torch
    dist = _dist()
    is_tensor = hasattr(local, "data_ptr")
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local if is_tensor else np.asarray(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    nccl = dist.get_backend() == "nccl"
    bounds = _counts(n_units, world, counts)
    if is_tensor:
        t = local.movedim(axis, 0).contiguous()
        if nccl and not t.is_cuda:
            t = t.cuda()
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.moveaxis(np.asarray(local), axis, 0)))
        if nccl:
            t = t.cuda()
    lo, hi = bounds[rank]
    assert t.shape[0] == hi - lo, "local shard has %d units, expected %d" % (t.shape[0], hi - lo)
    out = torch.empty((n_units,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    if len({b[1] - b[0] for b in bounds}) == 1:
        dist.all_gather_into_tensor(out, t)
    else:
        for r, (a, b) in enumerate(bounds):
            if b > a:
                view = out[a:b]
                if r == rank:
                    view.copy_(t)
                dist.broadcast(view, src=r)
    out = out.movedim(0, axis)
    if is_tensor:
        return out
    return out.cpu().numpy()


def _all_reduce_min(value):
    """Smallest value over the ranks (error codes are negative: any rank's failure wins)."""
    import torch
    dist = _dist()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return int(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())


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
    # one gather for the three result arrays: [3][R][prn]
    packed = np.stack([res["carrFreq"], res["codePhase"], res["peakMetric"]])
    full = gather_results(packed, nsat, axis=2)
    return dict(carrFreq=full[0], codePhase=full[1], peakMetric=full[2])


def track_sharded(recordings, rec_len, channel_sets, settings, stream=0, gather=True):
    """Tracking of this rank's recordings (the caller passes only the local shard: recordings live on
    the GPU that tracks them).  Returns ``(rc, out, ms_done)`` with ``rc`` the most severe status of ANY rank (so a
    short recording on one rank is seen by all).  ``gather=True``: ``out [R_total, C, 13, ms]`` and ``ms_done`` of
    all ranks in rank order; with device-resident recordings the result buffers stay on the GPUs and are gathered
    GPU to GPU.  ``gather=False``: every rank keeps (and may copy to its own host) its shard only -- SURVEY.md
    section 8(e) allows either.  Rows of idle (PRN 0) or stopped channels are zero, not uninitialised."""
    from .tracking import track_batch
    from ._native import TRACK_FIELDS
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    on_device = hasattr(recordings, "data_ptr") and recordings.is_cuda
    ms = int(settings.msToProcess)
    shape = (len(channel_sets), int(settings.numberOfChannels), len(TRACK_FIELDS), ms)
    if on_device:
        import torch
        out = torch.zeros(shape, dtype=torch.float64, device=recordings.device)
    else:
        out = np.zeros(shape, dtype=np.float64)
    rc, out, done = track_batch(recordings, rec_len, channel_sets, settings, out=out, stream=stream)
    if world == 1:
        return rc, out, done
    rc = _all_reduce_min(rc)
    if not gather:
        return rc, out, done
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n_local = torch.tensor([len(channel_sets)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    total = sum(counts)
    return rc, gather_results(out, total, axis=0, counts=counts), gather_results(done, total, axis=0, counts=counts)


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

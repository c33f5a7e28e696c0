"""Host-side multi-GPU logic on CPU: world_size-2 gloo.  The compute itself needs a GPU, so the local
shard results are synthesised; what is tested is the partition and the gather order."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from softgnss_python_b200 import dist as sd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_prn = 31                                      # uneven split on purpose
    lo, hi = sd.shard_range(n_prn, rank, world)
    full = np.arange(3 * n_prn, dtype=np.float64).reshape(3, n_prn)     # [recordings, PRN]
    got = sd.gather_results(full[:, lo:hi], n_prn, axis=1)
    trk = np.arange(4 * 2 * 13 * 5, dtype=np.float64).reshape(4, 2, 13, 5)  # 4 recordings total
    l2, h2 = sd.shard_range(4, rank, world)
    got2 = sd.gather_results(trk[l2:h2], 4, axis=0)
    # navigation solutions: the epoch count differs per rank (ragged second axis), padded with NaN / 0
    e_local = 3 if rank == 0 else 5
    sol = np.full((h2 - l2, e_local, 12), float(rank + 1))
    act = np.full((h2 - l2, e_local, 8), rank + 1, dtype=np.uint8)
    gs, ga = sd.gather_epochs(sol, 4), sd.gather_epochs(act, 4)
    ok3 = gs.shape == (4, 5, 12) and ga.shape == (4, 5, 8)
    for r2 in range(world):
        a, b = sd.shard_range(4, r2, world)
        e2 = 3 if r2 == 0 else 5
        ok3 = ok3 and np.all(gs[a:b, :e2] == r2 + 1) and np.all(np.isnan(gs[a:b, e2:]))
        ok3 = ok3 and np.all(ga[a:b, :e2] == r2 + 1) and np.all(ga[a:b, e2:] == 0)
    q.put((rank, np.array_equal(got, full), np.array_equal(got2, trk) and bool(ok3), (lo, hi)))
    dist.destroy_process_group()


def test_shard_ranges_partition():
    from softgnss_python_b200.dist import shard_range
    for n in (0, 1, 7, 32, 256):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_gather_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] and r[2] for r in res), res
    assert sorted(r[3] for r in res) == [(0, 16), (16, 31)]

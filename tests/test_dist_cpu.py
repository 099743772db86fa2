"""CPU tests (gloo, world_size 2) of the host-side logic of the multi-GPU path: strip partition,
scene sharding (every particle owned exactly once) and the max-over-ranks / sum-over-ranks
reductions bench.py uses.  No CUDA."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ptoy_b200 import partition_columns, scenes


def test_partition_columns_properties():
    for n, w in [(100, 1), (100, 2), (101, 4), (1880, 8), (16, 8)]:
        cb = partition_columns(n, w)
        assert cb[0] == 0 and cb[-1] == n and len(cb) == w + 1
        assert np.all(np.diff(cb) >= 2)
    w = np.zeros(100)
    w[60:] = 1.0                       # all the particles in the right 40 % of the domain
    cb = partition_columns(100, 4, weights=w)
    assert cb[1] >= 60 and np.all(np.diff(cb) >= 2)
    per = [w[cb[r]:cb[r + 1]].sum() for r in range(4)]
    assert max(per[1:]) <= 14


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.strip_scene(20000, rank, world)
    # ownership: the union over ranks is the global scene, no id twice
    ids = torch.zeros(sc.n_global, dtype=torch.int32)
    ids[torch.from_numpy(sc.ids.astype(np.int64))] = 1
    dist.all_reduce(ids)
    ok = bool((ids == 1).all())
    # the reductions of bench.py: time = max over ranks, particles/launches = sum
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(sc.n)], dtype=torch.float64)
    dist.all_reduce(n)
    # the NCCL id travels as bytes through a broadcast
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid[:] = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    if rank == 0:
        q.put((ok, float(t.item()), float(n.item()), sc.n_global, int(uid.sum())))
    dist.destroy_process_group()


def test_two_rank_host_logic_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, tmax, nsum, nglob, uidsum = q.get(timeout=10)
    assert ok and tmax == 2.0 and nsum == nglob and uidsum == sum(range(128))

"""Host-side multi-rank logic on CPU: shard bookkeeping and the reduction pattern of the
edge-partitioned assembly, world_size 2 over gloo (the GPU path uses NCCL with the same calls)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_tiles_exactly(pkg):
    from importlib import import_module
    par = import_module(pkg.__name__ + ".parallel")
    for n in (0, 1, 7, 4096, 100_000, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            sizes = []
            for r in range(world):
                lo, hi = par.shard_range(n, r, world)
                assert lo == prev and hi >= lo
                prev = hi
                sizes.append(hi - lo)
            assert prev == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_poses, n_lm, pkg_name, out):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    par = importlib.import_module(pkg_name + ".parallel")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every edge contributes to exactly one rank's partial landmark sums (edges follow their pose)
    rng = np.random.default_rng(0)
    edge_pose = rng.integers(0, n_poses, 5000)
    edge_lm = rng.integers(0, n_lm, 5000)
    contrib = rng.normal(size=(5000, 6))
    lo, hi = par.shard_range(n_poses, rank, world)
    mine = (edge_pose >= lo) & (edge_pose < hi)
    part = np.zeros((n_lm, 6))
    np.add.at(part, edge_lm[mine], contrib[mine])
    t = torch.from_numpy(part.reshape(-1).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    full = np.zeros((n_lm, 6))
    np.add.at(full, edge_lm, contrib)
    ok = np.allclose(t.numpy().reshape(n_lm, 6), full, rtol=1e-12, atol=1e-12)
    # time-like reduction used by bench.py: max over ranks
    tm = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    out[rank] = int(ok and tm.item() == world)
    dist.destroy_process_group()


def test_edge_partitioned_reduction_world2_gloo(pkg):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        out = m.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, 1000, 37, pkg.__name__, out)) for r in range(world)]
        [p.start() for p in procs]
        [p.join(120) for p in procs]
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: 1, 1: 1}


def _gather_worker(rank, world, port, pkg_name, out):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    par = importlib.import_module(pkg_name + ".parallel")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # what connect_peer_exchange gathers: fixed-size opaque handles and int32 landmark ranges, rank order
    handle = bytes([(rank * 37 + k) % 256 for k in range(64)])
    allh = par.gather_bytes(handle, world)
    rng = np.array([100 * rank, 100 * rank + 130], dtype=np.int32)
    allr = np.frombuffer(par.gather_bytes(rng.tobytes(), world), dtype=np.int32).reshape(world, 2)
    ok = len(allh) == 64 * world and all(allh[64 * r:64 * (r + 1)] == bytes([(r * 37 + k) % 256 for k in range(64)])
                                         for r in range(world))
    ok = ok and allr.tolist() == [[100 * r, 100 * r + 130] for r in range(world)]
    out[rank] = int(ok)
    dist.destroy_process_group()


def test_gather_bytes_world2_gloo(pkg):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        out = m.dict()
        procs = [ctx.Process(target=_gather_worker, args=(r, world, port, pkg.__name__, out)) for r in range(world)]
        [p.start() for p in procs]
        [p.join(120) for p in procs]
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: 1, 1: 1}

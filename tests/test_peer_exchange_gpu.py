"""Peer-memory exchange of the landmark part between pose-range shards (config 5 path) on >= 2 GPUs of one
node: identical to the NCCL all-reduce path and to the unsharded assembly.  Skipped on a 1-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import importlib
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    pkg = load_package()
    par = importlib.import_module(pkg.__name__ + ".parallel")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = pkg.Context(rank, stream=stream.cuda_stream)
    g = pkg.synth.c5_graph(n_poses=30_000, n_pairs=3_000)
    P, L = len(g.pose_ids), len(g.lm_ids)
    ctx.graph_load(g)
    ctx.graph_prepare_assembly_only()
    with torch.cuda.stream(stream):
        # reference 1: the whole graph on this rank
        ctx.graph_assemble_async(0, P)
        full = par.landmark_part_tensor(ctx, dev).clone()
        # reference 2: pose shards + NCCL all-reduce
        par.assemble_sharded(ctx, P, rank, world, device=dev)
        nccl = par.landmark_part_tensor(ctx, dev).clone()
    stream.synchronize()
    ranges = par.connect_peer_exchange(ctx, P, rank, world, device=dev)
    results = []
    with torch.cuda.stream(stream):
        for _ in range(5):   # several epochs: both parities, flag reuse
            par.landmark_part_tensor(ctx, dev).fill_(123.0)
            par.assemble_sharded(ctx, P, rank, world, device=dev, peer=True)
            results.append(par.landmark_part_tensor(ctx, dev).clone())
    stream.synchronize()
    err = ctx.xchg_error()
    scale = float(full.abs().max().item())
    ok = err == 0
    for t in results:
        ok = ok and bool(torch.equal(t, results[0]))                       # same bits every epoch
        ok = ok and float((t - nccl).abs().max().item()) <= 1e-12 * scale   # == all-reduce (sum order may differ)
        ok = ok and float((t - full).abs().max().item()) <= 1e-10 * scale   # == unsharded assembly
    # every rank holds the same bits
    mine = results[0].clone()
    other = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(other, mine)
    ok = ok and all(bool(torch.equal(o, mine)) for o in other)
    out[rank] = (int(ok), int(err), ranges.tolist(), scale)
    ctx.close()
    dist.destroy_process_group()


def test_peer_exchange_equals_allreduce_and_unsharded():
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs on one node")
    world = min(torch.cuda.device_count(), 4)
    port = _free_port()
    mctx = mp.get_context("spawn")
    with mctx.Manager() as m:
        out = m.dict()
        procs = [mctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        [p.start() for p in procs]
        [p.join(300) for p in procs]
        for p in procs:
            if p.is_alive():
                p.kill()
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        res = dict(out)
        assert all(res[r][0] == 1 for r in range(world)), res

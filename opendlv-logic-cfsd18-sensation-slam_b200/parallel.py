"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

Only where the path shards:
  * independent graph replicas (config 3) and observation batches (config 4): contiguous ranges per
    rank, no collective on the data path;
  * a single large graph (config 5): edges partitioned by pose range; pose blocks and all
    off-diagonal blocks have a unique owner, only the landmark part of the normal equations (the
    first 6L doubles of the system array: b_lm 2L | H_lm 4L) is summed across ranks.
A single small graph (configs 1-2) does not shard: replicas only.
"""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of n units for `rank` of `world` (sizes differ by at most 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DeviceArray:
    """Zero-copy view of a raw device pointer for torch (``torch.as_tensor(DeviceArray(...))``)."""

    def __init__(self, ptr: int, n: int, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def landmark_part_tensor(ctx, device):
    """torch view (no copy) of the part of the assembled system that pose-range shards must reduce."""
    import torch
    ptr, n = ctx.graph_system_dev(0)
    if n == 0:
        return torch.empty(0, dtype=torch.float64, device=device)
    return torch.as_tensor(DeviceArray(ptr, n), device=device)


def gather_bytes(payload: bytes, world: int, device=None, group=None):
    """All-gather of equal-length byte strings over the process group (rank order)."""
    import torch
    import torch.distributed as dist
    t = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)


def connect_peer_exchange(ctx, n_poses: int, rank: int, world: int, device=None, group=None):
    """Collective set-up of the peer-memory exchange of the landmark part (one process per GPU of one
    node): every rank computes the landmark range its pose shard touches, allocates its exchange region
    and the CUDA IPC handles and ranges are gathered over the ranks.  Returns the gathered ranges."""
    import numpy as np
    lo, hi = shard_range(n_poses, rank, world)
    l0, l1 = ctx.graph_shard_landmarks(lo, hi)
    mine = np.array([l0, l1], dtype=np.int32)
    ranges = np.frombuffer(gather_bytes(mine.tobytes(), world, device, group), dtype=np.int32).reshape(world, 2)
    cap = int(max(1, (ranges[:, 1] - ranges[:, 0]).max()))
    handle = ctx.xchg_create(world, rank, cap)
    handles = gather_bytes(handle, world, device, group)
    ctx.xchg_connect(handles, ranges)
    return ranges


def assemble_sharded(ctx, n_poses: int, rank: int, world: int, device=None, group=None, peer=False):
    """Edge-partitioned assembly of one large graph: this rank linearises the edges of its pose range,
    then the landmark diagonal blocks and landmark rhs are summed over ranks -- with an NCCL all-reduce
    over NVLink, or (peer=True, after connect_peer_exchange) by the landmark kernel's own stores into
    every rank's memory followed by a rank-ordered sum.  The caller's stream must be the context's
    stream.  Returns the pose range."""
    import torch.distributed as dist
    lo, hi = shard_range(n_poses, rank, world)
    if world > 1 and peer:
        ctx.graph_assemble_exchange_async(lo, hi)
        return lo, hi
    ctx.graph_assemble_async(lo, hi)
    if world > 1:
        t = landmark_part_tensor(ctx, device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return lo, hi

"""A/B of the pose-centred assembly kernel variants (SLAM_B200_ASM_VARIANT = 10*depth + minBlocksPerSM,
0 = original kernel) on config 5 (1M-pose graph) and config 3 (replicas of the 1-lap graph):
linearise + assemble alone, L2 flushed before every launch pair, median of `reps`; V of every variant
is compared with the original kernel's."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
par = __import__(pkg.__name__ + ".parallel", fromlist=["x"])
synth = pkg.synth
VARIANTS = [int(v) for v in os.environ.get("ASM_AB_VARIANTS", "0,14,24").split(",")]
GROUPS = [int(v) for v in os.environ.get("ASM_AB_GROUPS", "32,16,8").split(",")]  # lanes per landmark
REPS = int(os.environ.get("ASM_AB_REPS", "15"))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def measure(ctx, P, label, nV, R):
    ptr, _ = ctx.graph_system_dev(0)
    V = torch.as_tensor(par.DeviceArray(ptr, nV * R), device=dev)
    ref = None
    for var, grp in [(v, g_) for v in VARIANTS for g_ in GROUPS]:
        os.environ["SLAM_B200_ASM_VARIANT"] = str(var)
        os.environ["SLAM_B200_LM_GROUP"] = str(grp)
        ts = []
        with torch.cuda.stream(stream):
            for _ in range(3):
                ctx.graph_assemble_async(0, P)
            for _ in range(REPS):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.graph_assemble_async(0, P)
                e1.record(stream)
                stream.synchronize()
                ts.append(e0.elapsed_time(e1))
        got = V.clone()
        if ref is None:
            ref = got
            diff = 0.0
        else:
            diff = float((got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-300))
        print("%s pose-kernel variant %2d, %2d lanes per landmark: median %.4f ms  min %.4f ms  max|dV|/max|V| vs first = %.3g"
              % (label, var, grp, float(np.median(ts)), float(np.min(ts)), diff), flush=True)
        del got


if "c5" in os.environ.get("ASM_AB_WORKLOADS", "c5,c3"):
    ctx = pkg.Context(0, stream=stream.cuda_stream)
    g = synth.c5_graph()
    ctx.graph_load(g)
    ctx.graph_prepare_assembly_only()
    st = ctx.graph_stats()
    measure(ctx, len(g.pose_ids), "c5", int(st["nV"]), 1)
    ctx.close()
if "c2" in os.environ.get("ASM_AB_WORKLOADS", "c5,c3"):
    ctx = pkg.Context(0, stream=stream.cuda_stream)
    g = synth.c2_graph()
    ctx.graph_load(g)
    ctx.graph_prepare_assembly_only()
    st = ctx.graph_stats()
    measure(ctx, len(g.pose_ids), "c2", int(st["nV"]), 1)
    ctx.close()
if "c3" in os.environ.get("ASM_AB_WORKLOADS", "c5,c3"):
    R = int(os.environ.get("ASM_AB_REPLICAS", "2048"))
    ctx = pkg.Context(0, stream=stream.cuda_stream)
    g = synth.graph_from_drive(synth.trackdrive(1))
    b = synth.perturb_replicas(g, R, seed=18, first=0)
    ctx.graph_load(g)
    ctx.batch_upload(b[0], b[1], b[3], b[2])
    st = ctx.graph_stats()
    measure(ctx, len(g.pose_ids), "c3x%d" % R, int(st["nV"]), R)
    ctx.close()

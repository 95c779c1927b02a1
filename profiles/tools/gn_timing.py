"""GN iteration time of one graph (c1 | c2) for the nested-dissection leaf size given by
SLAM_B200_ND_LEAF: usage  gn_timing.py c1"""
import os
import sys
import time
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
synth = pkg.synth
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
g = synth.graph_from_drive(synth.trackdrive(1)) if which == "c1" else synth.c2_graph()
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
ctx.graph_load(g)
ctx.graph_prepare()
ctx.graph_snapshot()
st = ctx.graph_stats()
with torch.cuda.stream(stream):
    for _ in range(3):
        ctx.graph_restore_async(); ctx.graph_iterate_async(10)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        ctx.graph_restore_async(); ctx.graph_iterate_async(10)
    e1.record(stream)
stream.synchronize()
ms = e0.elapsed_time(e1) / 100
print("%s leaf=%s: %.1f us per GN iteration (%.0f it/s), levels %d, fronts %d, nnzL %d, max front %d, symbolic %.1f ms"
      % (which, os.environ.get("SLAM_B200_ND_LEAF", "default"), ms * 1e3, 1e3 / ms, st["n_levels"], st["n_fronts"],
         st["nnz_L"], st["max_front"], st["symbolic_seconds"] * 1e3))

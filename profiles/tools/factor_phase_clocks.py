"""Phase breakdown (SM cycles) of the root front's factor CTA on the 10-lap trackdrive graph.
Run with SLAM_B200_PHASE_CLOCKS=1 SLAM_B200_NO_CUDA_GRAPH=1."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
ctx = pkg.Context(0)
ctx.graph_load(pkg.synth.c2_graph())
ctx.graph_prepare()
ctx.graph_iterate_async(3)
ctx.sync()
out = (C.c_longlong * 10)()
rc = ctx.L.slam_b200_debug_phase_clocks(ctx.h, out)
v = list(out)
print("rc", rc, "s", v[7], "fs", v[8], "children", v[9])
names = ["zero", "scatter H", "extend-add", "LDL^T panels", "fused forward", "write L + U"]
tot = v[6] - v[0]
for k, nm in enumerate(names):
    print("%-14s %8d cycles %5.1f%%" % (nm, v[k + 1] - v[k], 100.0 * (v[k + 1] - v[k]) / tot))
print("total %d cycles" % tot)
out8 = (C.c_longlong * 8)()
ctx.L.slam_b200_debug_panel_clocks.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
rc = ctx.L.slam_b200_debug_panel_clocks(ctx.h, out8)
w = list(out8)
print("panel loop of that CTA, thread 0 (warp 0), cycles summed over its panels:")
for nm, val in zip(["(1a) triangle, warp 0 (gen 2: look-ahead factorise+publish)", "wait at barrier 1 (gen 2: look-ahead update of the triangle)", "(1b) row elimination", "wait at barrier 2",
                    "(2) trailing update", "wait at barrier 3", "forward: warp-0 triangle solve", "forward: rest of panel step"], w):
    print("  %-64s %9.0f" % (nm, val))

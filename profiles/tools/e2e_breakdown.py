"""Where a COLD optimise call goes (config 2): graph_load, graph_prepare (host structure pass + symbolic
analysis + upload), 10 GN iterations, estimates read-back -- timed separately on the host clock."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
g = pkg.synth.c2_graph()
ctx = pkg.Context(0)
acc = {}
for rep in range(8):
    t = [time.perf_counter()]
    ctx.graph_load(g); t.append(time.perf_counter())
    ctx.graph_prepare(); t.append(time.perf_counter())
    n, chi2 = ctx.graph_optimize(10); t.append(time.perf_counter())
    pe, le = ctx.graph_get_estimates(); t.append(time.perf_counter())
    if rep >= 3:
        for k, name in enumerate(("graph_load", "graph_prepare", "graph_optimize(10)", "get_estimates")):
            acc.setdefault(name, []).append((t[k + 1] - t[k]) * 1e3)
st = ctx.graph_stats()
tot = 0.0
for k, v in acc.items():
    v.sort(); tot += v[len(v) // 2]
    print("%-20s median %7.3f ms" % (k, v[len(v) // 2]))
print("%-20s        %7.3f ms  -> %.0f GN it/s" % ("total", tot, 10 / tot * 1e3))
print("inside graph_prepare (last call): structure %.3f ms, symbolic %.3f ms, launch lists %.3f ms, upload %.3f ms" % (
    st["structure_seconds"] * 1e3, st["symbolic_seconds"] * 1e3, st["launch_lists_seconds"] * 1e3, st["upload_seconds"] * 1e3))

"""Phase breakdown (SM cycles summed over all warps) of the warp-per-front factor kernel on the batched
Monte-Carlo configuration (C3).  Run with SLAM_B200_PHASE_CLOCKS=1."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
synth = pkg.synth
g = synth.graph_from_drive(synth.trackdrive(1))
R = 1024
b = synth.perturb_replicas(g, R, seed=18, first=0)
ctx = pkg.Context(0)
ctx.graph_load(g)
ctx.batch_upload(b[0], b[1], b[3], b[2])
ctx.batch_iterate_async(2)
ctx.sync()
out = (C.c_longlong * 8)()
ctx.L.slam_b200_debug_tiny_clocks.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
rc = ctx.L.slam_b200_debug_tiny_clocks(ctx.h, out)
v = list(out)
names = ["zero", "scatter H", "extend-add", "LDL^T panels", "fused forward", "write L + U"]
tot = sum(v[:6])
if rc != 0 or tot == 0:
    print("phase clocks not enabled (SLAM_B200_PHASE_CLOCKS=1); rc", rc)
    sys.exit(0)
print("rc", rc, "fronts", v[6], "mean cycles per front %.0f" % (tot / max(v[6], 1)))
for k, nm in enumerate(names):
    print("%-14s %6.0f cycles/front %5.1f%%" % (nm, v[k] / max(v[6], 1), 100.0 * v[k] / tot))

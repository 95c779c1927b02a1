"""Front timeline of one Gauss-Newton iteration on the 10-lap trackdrive graph (SLAM_B200_TIMELINE=1): per level of the
assembly tree when its factor / backward CTAs entered, passed their dependency wait and ended (us since the first stamp),
how long a front took after the wait, and the gap between the end of one level and the release of the next."""
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

os.environ["SLAM_B200_TIMELINE"] = "1"
pkg = load_package()
ctx = pkg.Context(0)
laps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx.graph_load(pkg.synth.c2_graph() if laps == 10 else pkg.synth.graph_from_drive(pkg.synth.trackdrive(laps)))
ctx.graph_prepare()
ctx.graph_iterate_async(6)
ctx.sync()
L = ctx.L
L.slam_b200_debug_timeline.restype = C.c_long
L.slam_b200_debug_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_long]
n = L.slam_b200_debug_timeline(ctx.h, None, 0)
buf = (C.c_longlong * n)()
L.slam_b200_debug_timeline(ctx.h, buf, n)
raw = np.array(buf, dtype=np.int64).reshape(-1, 12)
tl = raw.astype(np.float64) / 1e3
L.slam_b200_graph_export_symbolic.restype = C.c_long
L.slam_b200_graph_export_symbolic.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_long]


def export(what):
    m = L.slam_b200_graph_export_symbolic(ctx.h, what, None, 0)
    a = np.zeros(m, dtype=np.int32)
    L.slam_b200_graph_export_symbolic(ctx.h, what, a.ctypes.data_as(C.POINTER(C.c_int32)), m)
    return a


level_ptr, npiv, nupd = export(3), export(5), export(6)
t0 = tl[tl > 0].min()
tl -= t0
print("F = factor launch(es) of a level, B = backward; times in us since the first stamp; phases = medians over the level's fronts")
prev_end = 0.0
for l in range(len(level_ptr) - 1):
    a, b = level_ptr[l], level_ptr[l + 1]
    t = tl[a:b]
    dur = t[:, 5] - t[:, 1]
    k = int(np.argmax(dur))
    ph = [np.median(t[:, i + 1] - t[:, i]) for i in range(1, 5)]
    phk = [t[k, i + 1] - t[k, i] for i in range(1, 5)]
    print("F L%-2d %5d | entry %6.1f released %6.1f..%6.1f (gap %4.1f) end %6.1f | after wait: median %5.1f max %5.1f | rhs+maps %4.1f children %4.1f panels %4.1f write %4.1f | slowest s=%d u=%d: %4.1f %4.1f %4.1f %4.1f"
          % (l, b - a, t[:, 0].min(), t[:, 1].min(), t[:, 1].max(), t[:, 1].min() - prev_end, t[:, 5].max(), np.median(dur), dur.max(),
             ph[0], ph[1], ph[2], ph[3], npiv[a + k], nupd[a + k], phk[0], phk[1], phk[2], phk[3]))
    prev_end = t[:, 5].max()
t_factor_end = prev_end
for l in range(len(level_ptr) - 2, -1, -1):
    a, b = level_ptr[l], level_ptr[l + 1]
    t = tl[a:b]
    if raw[a:b, 6].min() == 0:
        print("B L%-2d %5d | solved inside the factor kernel (roots)" % (l, b - a))
        continue
    dur = t[:, 8] - t[:, 7]
    print("B L%-2d %5d | entry %6.1f released %6.1f..%6.1f (gap %4.1f) end %6.1f | after wait: median %5.1f max %5.1f"
          % (l, b - a, t[:, 6].min(), t[:, 7].min(), t[:, 7].max(), t[:, 7].min() - prev_end, t[:, 8].max(), np.median(dur), dur.max()))
    prev_end = t[:, 8].max()
print("factor %.1f us, backward %.1f us" % (t_factor_end, prev_end - t_factor_end))

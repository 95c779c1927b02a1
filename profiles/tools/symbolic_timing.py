"""Times the host symbolic phase (nested dissection + region-parallel minimum degree + fronts) on the
10-lap trackdrive pattern.  SLAM_B200_SYM_DEBUG=1 prints the per-stage breakdown to stderr."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from conftest import load_pkg  # noqa: E402
import mf_emul  # noqa: E402

pkg = load_pkg()
g = pkg.synth.c2_graph()
ids, dims, pa, pb = mf_emul.block_pattern(g)
L = pkg.capi.lib()
capi = pkg.capi
d = capi._i32(dims); a = capi._i32(pa); b = capi._i32(pb)
for rep in range(3):
    h = C.c_void_p(L.slam_b200_symbolic_create(len(d), capi._ip(d), len(a), capi._ip(a), capi._ip(b), 1024))
    print("total %.4f s  nd %.4f  md %.4f   nnz(L) %d  flops %.4g  max front %d" % (
        L.slam_b200_symbolic_stat(h, 3), L.slam_b200_symbolic_stat(h, 5), L.slam_b200_symbolic_stat(h, 6),
        L.slam_b200_symbolic_stat(h, 0), L.slam_b200_symbolic_stat(h, 1), L.slam_b200_symbolic_stat(h, 2)))
    L.slam_b200_symbolic_destroy(h)

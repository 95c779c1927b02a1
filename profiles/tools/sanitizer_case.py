"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): association frames, bulk
association, one graph optimisation (CTA and warp kernels), a small replica batch."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

SMALL = os.environ.get("SLAM_B200_SANITIZER_SMALL") == "1"   # the pytest wrapper: minutes under racecheck otherwise
pkg = load_package()
synth = pkg.synth
ctx = pkg.Context(0)
trk = synth.ellipse_track()
d = synth.simulate_drive(trk, 30 if SMALL else 80, s_step=trk.length / 1000, seed=7)
cci = lc = 0
for fr, p in zip(d.frames, d.poses_noisy):
    r = ctx.assoc_map_frame(fr, p, 1.2, 50.0, cci, lc)
    cci, lc = r["cci"], r["loop_closing"]
print("map", ctx.map_size())
r = ctx.assoc_localize_frame(d.frames[10], d.poses_noisy[10], 1.2, cci)
f = synth.cone_field(n_map=4000 if SMALL else 20000, n_obs=600 if SMALL else 3000, seed=4)
ctx.map_clear(); ctx.map_append(f.map_x, f.map_y, f.map_type)
a = ctx.assoc_bulk(f.frame, f.pose, 1.2, 0, pkg.capi.ALGO_GRID).copy()
b = ctx.assoc_bulk(f.frame, f.pose, 1.2, 0, pkg.capi.ALGO_BRUTE).copy()
assert np.array_equal(a, b)
g = synth.graph_from_drive(synth.simulate_drive(trk, 150 if SMALL else 400, s_step=trk.length / 1000, seed=3))
ctx.graph_load(g)
print("optimize", ctx.graph_optimize(2 if SMALL else 3))
pe, le, ez, oz = synth.perturb_replicas(g, 8 if SMALL else 40, seed=18)
out = ctx.graph_optimize_batch(pe, le, oz, ez, iters=2)
print("batch", out[3])
ctx.close()

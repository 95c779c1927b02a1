"""Small end-to-end case run in its own process under a memory checker: association frames (mapping and
localiser), grid and brute-force bulk association, one graph optimisation through the CTA-per-front and
the warp-per-front kernels, a small replica batch, a sharded assembly -- every result compared with the
CPU oracle, so a poisoned (out-of-bounds / uninitialised) read shows up as a mismatch.

  compute-sanitizer --tool memcheck|racecheck python tests/gpu_case_sanitizer.py      (where the tool is open)
  SLAM_B200_GUARD=1 python tests/gpu_case_sanitizer.py                                 (the library's own guard bands)

Prints one JSON line: guard-band verdict + a hash of every result (two runs must hash equal: the kernels have
no atomics on value paths, so a data race is the only source of run-to-run differences)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from __graft_entry__ import load_package  # noqa: E402
from oracle import oracle  # noqa: E402

SMALL = os.environ.get("SLAM_B200_SANITIZER_SMALL") == "1"   # minutes under racecheck otherwise
pkg = load_package()
orc = oracle.load("best")
synth = pkg.synth
ctx = pkg.Context(0)
H = hashlib.sha256()


def eat(*arrays):
    for a in arrays:
        H.update(np.ascontiguousarray(a).tobytes())


def close(a, b, tol=1e-6):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape and np.all(np.isfinite(a)), "non-finite or misshapen result"
    assert np.max(np.abs(a - b)) <= tol * max(1.0, float(np.max(np.abs(b)))), float(np.max(np.abs(a - b)))


trk = synth.ellipse_track()
d = synth.simulate_drive(trk, 30 if SMALL else 80, s_step=trk.length / 1000, seed=7)
cap = 4096
mx = np.zeros(cap); my = np.zeros(cap); mt = np.zeros(cap, dtype=np.int32)
M = occi = olc = cci = lc = 0
for fr, p in zip(d.frames, d.poses_noisy):
    o = orc.assoc_map_frame(fr, p, 1.2, 50.0, mx, my, mt, M, occi, olc)
    r = ctx.assoc_map_frame(fr, p, 1.2, 50.0, cci, lc)
    assert np.array_equal(r["idx"], o["idx"]) and np.array_equal(r["status"], o["status"]), "mapping association differs"
    M, occi, olc = o["M"], o["cci"], o["loop_closing"]
    cci, lc = r["cci"], r["loop_closing"]
    eat(r["idx"], r["status"])
assert ctx.map_size() == M
print("map", M)
r = ctx.assoc_localize_frame(d.frames[10], d.poses_noisy[10], 1.2, cci)
eat(r["idx"])
f = synth.cone_field(n_map=4000 if SMALL else 20000, n_obs=600 if SMALL else 3000, seed=4)
ctx.map_clear(); ctx.map_append(f.map_x, f.map_y, f.map_type)
a = ctx.assoc_bulk(f.frame, f.pose, 1.2, 0, pkg.capi.ALGO_GRID).copy()
b = ctx.assoc_bulk(f.frame, f.pose, 1.2, 0, pkg.capi.ALGO_BRUTE).copy()
o = orc.assoc_match_only(np.asfortranarray(f.frame), f.pose, 1.2, 0, f.map_x, f.map_y, f.map_type)
assert np.array_equal(a, b) and np.array_equal(a, o["idx"]), "bulk association differs"
eat(a)
g = synth.graph_from_drive(synth.simulate_drive(trk, 150 if SMALL else 400, s_step=trk.length / 1000, seed=3))
ctx.graph_load(g)
it = 2 if SMALL else 3
n, chi2 = ctx.graph_optimize(it)
G = orc.graph_from_soa(g)
no, chi2o = G.optimize(it)
pe, le = ctx.graph_get_estimates()
po, lo = G.estimates(g)
assert n == no == it
close(chi2, chi2o, 1e-8); close(pe, po); close(le, lo)
eat(chi2, pe, le)
print("optimize", n)
R = 8 if SMALL else 40
bpe, ble, bez, boz = synth.perturb_replicas(g, R, seed=18)
out = ctx.graph_optimize_batch(bpe, ble, boz, bez, iters=2)
import copy  # noqa: E402
for q in (0, R - 1):   # first and last replica against the oracle
    gq = copy.copy(g)
    gq.pose_est, gq.lm_est, gq.el_z, gq.eo_z = bpe[q], ble[q], bez[q], boz[q]
    Gq = orc.graph_from_soa(gq)
    Gq.optimize(2)
    pq, lq = Gq.estimates(gq)
    close(out[0][q], pq); close(out[1][q], lq)
assert np.all(np.isfinite(out[0])) and np.all(np.isfinite(out[1])) and np.all(np.asarray(out[3]) == 2)
eat(out[0], out[1], out[2])
print("batch", out[3])
# sharded assembly: two pose ranges
ctx.graph_load(g)
ctx.graph_prepare_assembly_only()
P = len(g.pose_ids)
for p0, p1 in ((0, P // 2), (P // 2, P)):
    ctx.graph_assemble_async(p0, p1)
ctx.sync()
bad, narr = ctx.debug_guard_check()
verdict = {"guard_mode": bad >= 0, "guard_bytes_changed": int(bad), "arrays_checked": int(narr), "hash": H.hexdigest()}
print(json.dumps(verdict))
ctx.close()
sys.exit(0 if bad <= 0 else 4)

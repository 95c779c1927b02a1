"""Generates the committed golden fixtures from the oracle (run in the build container, where
oracle/_ref -- the build against the reference's vendored Eigen -- exists):

    python tests/golden/make_golden.py

known_answers.json : survey-derived known answers of the conversion maths (g++ 13.3 /
                     glibc, -O2, no FMA), re-derived here by calling the oracle.
c1_replay.npz      : configuration C1 (trackdrive loop, seed 18) replayed through the restated Slam
                     back half: association record of every frame, final map, chi2 per GN iteration.
c1_graph_opt.npz   : the C1 graph with ground-truth association optimised 10 iterations: chi2 per
                     iteration and final estimates.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from conftest import load_pkg  # noqa: E402
from oracle import oracle  # noqa: E402

pkg = load_pkg()
synth = pkg.synth
o = oracle.load("best")
print("oracle kind:", o.kind)

ka = {"pi_ref": repr(o.lib.orc_pi_ref()), "spherical2cartesian": [], "cone_to_global": []}
for args in [(10, 0, 5), (-35.5, 0, 12.25), (80, 1, 3), (0, 0, 5)]:
    out = o.spherical2cartesian(*args)
    ka["spherical2cartesian"].append({"in": list(args), "out": [None if np.isnan(v) else float(v) for v in out]})
for obs in [(10, 0, 5, 1), (-35.5, 0, 12.25, 2), (80, 1, 3, 1)]:
    out = o.cone_to_global([3, -2, 0.7], obs)
    ka["cone_to_global"].append({"pose": [3, -2, 0.7], "obs": list(obs), "out": [float(out[0]), float(out[1])]})
json.dump(ka, open(os.path.join(HERE, "known_answers.json"), "w"), indent=1)

d = synth.trackdrive(1)
s = o.slam(synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
idx_all, st_all, rcs = [], [], []
for fr, p in zip(d.frames, d.poses_noisy):
    rc, idx, st = s.perform(fr, p)
    idx_all.append(idx); st_all.append(st); rcs.append(rc)
mx, my, mt = s.map()
np.savez_compressed(os.path.join(HERE, "c1_replay.npz"), idx=np.concatenate(idx_all), status=np.concatenate(st_all),
                    rc=np.array(rcs), map_x=mx, map_y=my, map_type=mt, chi2=s.chi2_log(),
                    cci=s.state()["current_cone_index"])
print("c1 replay:", s.state())

g = synth.graph_from_drive(d)
G = o.graph_from_soa(g)
n, chi2 = G.optimize(10)
pe, le = G.estimates(g)
np.savez_compressed(os.path.join(HERE, "c1_graph_opt.npz"), chi2=chi2, pose_est=pe, lm_est=le, iters=n)
print("c1 graph:", n, chi2[-1])

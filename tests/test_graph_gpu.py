"""Gauss-Newton step on the B200 through the C ABI vs the CPU oracle (fp64): assembled H and b,
chi2 per iteration, optimised poses and landmarks within 1e-6 relative (north_star tolerance)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import small_graph

pytestmark = pytest.mark.gpu

RTOL = 1e-6   # north_star: optimised poses and landmarks within 1e-6 relative
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel_close(a, b, rtol=RTOL):
    scale = max(1.0, float(np.max(np.abs(b))))
    return np.max(np.abs(a - b)) <= rtol * scale


def _to_dense_upper(s):
    n = s["n"]
    return sp.csc_matrix((s["Ax"], s["Ai"], s["Ap"]), shape=(n, n)).toarray()


@pytest.mark.parametrize("n_poses", [40, 300])
def test_assembled_system_matches_oracle(ctx, orc, synth, n_poses):
    g = small_graph(synth, n_poses)
    ctx.graph_load(g)
    d = ctx.graph_export_system()
    o = orc.graph_from_soa(g).build_system()
    assert d["n"] == o["n"]
    assert np.array_equal(d["Ap"], o["Ap"]) and np.array_equal(d["Ai"], o["Ai"])
    assert np.allclose(d["Ax"], o["Ax"], rtol=1e-10, atol=1e-12)
    assert np.allclose(d["b"], o["b"], rtol=1e-9, atol=1e-11)
    assert ctx.graph_chi2() == pytest.approx(o["chi2"], rel=1e-11)


def test_assembled_system_general_edges(ctx, orc, synth):
    """Duplicate pose-landmark edges (the doubled first-cone edge, slam.cpp:554-592), reversed and
    repeated pose-pose edges, anisotropic information, landmark ids above pose ids."""
    rng = np.random.default_rng(5)
    def build(G):
        for k in range(6):
            G.add_pose(1000 + k, *(rng_p[k]))
        for k in range(4):
            G.add_landmark(k if k < 2 else 5000 + k, *(rng_l[k]))
        for (a, b, z, info) in eo:
            G.add_edge_se2(a, b, z, info)
        for (p, l, z, info) in el:
            G.add_edge_se2_xy(p, l, z, info)
        G.set_fixed(1000, True)
        G.set_fixed(0, True)
    rng_p = rng.normal(size=(6, 3)); rng_l = rng.normal(size=(4, 2)) * 3
    def spd(n):
        a = rng.normal(size=(n, n)); return (a @ a.T + n * np.eye(n)).ravel()
    eo = [(1000, 1001, rng.normal(size=3), spd(3)), (1001, 1002, rng.normal(size=3), spd(3)),
          (1002, 1001, rng.normal(size=3), spd(3)), (1001, 1002, rng.normal(size=3), spd(3)),
          (1003, 1002, rng.normal(size=3), spd(3)), (1003, 1004, rng.normal(size=3), spd(3)),
          (1005, 1000, rng.normal(size=3), spd(3)), (1004, 1005, rng.normal(size=3), spd(3))]
    lms = [0, 1, 5002, 5003]
    el = [(1000 + int(rng.integers(0, 6)), lms[int(rng.integers(0, 4))], rng.normal(size=2), spd(2)) for _ in range(20)]
    el += [el[0], el[3], (1000, 0, rng.normal(size=2), spd(2))]   # duplicates + an all-fixed (inactive) edge

    class Dev:
        def add_pose(self, *a): ctx.graph_add_pose(*a)
        def add_landmark(self, *a): ctx.graph_add_landmark(*a)
        def add_edge_se2(self, *a): ctx.graph_add_edge_se2(*a)
        def add_edge_se2_xy(self, *a): ctx.graph_add_edge_se2_xy(*a)
        def set_fixed(self, *a): ctx.graph_set_fixed(*a)
    ctx.graph_clear()
    build(Dev())
    G = orc.graph(); build(G)
    d = ctx.graph_export_system(); o = G.build_system()
    assert d["n"] == o["n"] and np.array_equal(d["Ai"], o["Ai"])
    assert np.allclose(d["Ax"], o["Ax"], rtol=1e-10, atol=1e-12)
    assert np.allclose(d["b"], o["b"], rtol=1e-10, atol=1e-12)
    assert ctx.graph_chi2() == pytest.approx(o["chi2"], rel=1e-11)
    n1, c1 = ctx.graph_optimize(5); n2, c2 = G.optimize(5)
    assert n1 == n2 == 5 and np.allclose(c1, c2, rtol=1e-8)
    for vid in [1001, 1003, 1005, 1, 5002, 5003]:
        assert np.allclose(ctx.graph_get_vertex(vid), G.get_vertex(vid), rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("n_poses", [40, 300])
def test_optimize_small_graph(ctx, orc, synth, n_poses):
    g = small_graph(synth, n_poses)
    ctx.graph_load(g)
    n, chi2 = ctx.graph_optimize(10)
    G = orc.graph_from_soa(g)
    no, chi2o = G.optimize(10)
    assert n == no == 10
    assert np.allclose(chi2, chi2o, rtol=1e-8)
    pe, le = ctx.graph_get_estimates(); po, lo = G.estimates(g)
    assert _rel_close(pe, po) and _rel_close(le, lo)
    # fixed gauge untouched (slam.cpp:464-474)
    assert np.array_equal(pe[:2], g.pose_est[:2]) and np.array_equal(le[:2], g.lm_est[:2])


def test_optimize_c1_graph_golden(ctx, c1_graph):
    gold = np.load(os.path.join(GOLD, "c1_graph_opt.npz"))
    ctx.graph_load(c1_graph)
    n, chi2 = ctx.graph_optimize(10)
    assert n == int(gold["iters"]) == 10
    assert np.allclose(chi2, gold["chi2"], rtol=1e-8)
    pe, le = ctx.graph_get_estimates()
    assert _rel_close(pe, gold["pose_est"]) and _rel_close(le, gold["lm_est"])
    st = ctx.graph_stats()
    assert st["n"] == 3590 and st["n_levels"] < 40


def test_optimize_incremental_api_equals_bulk_load(ctx, synth):
    g = small_graph(synth, 60)
    ctx.graph_load(g)
    n, chi2 = ctx.graph_optimize(4)
    pe, le = ctx.graph_get_estimates()
    ctx.graph_clear()
    for i, vid in enumerate(g.lm_ids):
        ctx.graph_add_landmark(vid, *g.lm_est[i])
    for i, vid in enumerate(g.pose_ids):
        ctx.graph_add_pose(vid, *g.pose_est[i])
        if i > 0:
            ctx.graph_add_edge_se2(g.eo_from[i - 1], g.eo_to[i - 1], g.eo_z[i - 1], g.eo_info[i - 1])
    for e in range(len(g.el_pose)):
        ctx.graph_add_edge_se2_xy(g.el_pose[e], g.el_lm[e], g.el_z[e], g.el_info[e])
    for vid in g.fixed_ids:
        ctx.graph_set_fixed(vid, True)
    n2, chi22 = ctx.graph_optimize(4)
    pe2, le2 = ctx.graph_get_estimates()
    assert n == n2 == 4 and np.array_equal(chi2, chi22)   # deterministic assembly: bit-identical
    assert np.array_equal(pe, pe2) and np.array_equal(le, le2)


def test_odometry_measurement_matches_oracle(ctx, orc):
    rng = np.random.default_rng(1)
    ctx.graph_clear(); G = orc.graph()
    prev = np.array([1.0, 2.0, 3.0]); ctx.graph_add_pose(1000, *prev); G.add_pose(1000, *prev)
    info = np.eye(3).ravel() * 5
    for k in range(1, 6):
        cur = prev + rng.normal(size=3) * np.array([1, 1, 2.5])
        ctx.graph_add_pose(1000 + k, *cur); G.add_pose(1000 + k, *cur)
        ctx.graph_add_odometry(1000 + k - 1, 1000 + k, cur, info); G.add_odometry(1000 + k - 1, 1000 + k, cur, info)
        prev = cur
    ctx.graph_set_fixed(1000, True); G.set_fixed(1000, True)
    # measurement = prev^-1 * cur makes every edge error zero at the initial estimate
    assert ctx.graph_chi2() < 1e-25 and G.chi2() < 1e-25
    d = ctx.graph_export_system(); o = G.build_system()
    assert np.allclose(d["Ax"], o["Ax"], rtol=1e-10, atol=1e-12)


def test_return_codes(ctx):
    ctx.graph_clear()
    assert ctx.graph_optimize_rc(3)[0] == -1                     # empty graph
    ctx.graph_add_pose(1000, 0, 0, 0); ctx.graph_add_pose(1001, 1, 0, 0)
    ctx.graph_add_edge_se2(1000, 1001, [1, 0, 0], np.eye(3).ravel() * 5)
    ctx.graph_set_fixed(1000); ctx.graph_set_fixed(1001)
    assert ctx.graph_optimize_rc(3)[0] == -1                     # every edge inactive
    ctx.graph_set_fixed(1001, False)
    ctx.graph_add_pose(1002, 2, 0, 0)                            # free vertex without edges stays out
    assert ctx.graph_optimize_rc(3)[0] == 3
    ctx.graph_clear()
    ctx.graph_add_pose(1000, 0, 0, 0); ctx.graph_add_pose(1001, 1, 0, 0)
    ctx.graph_add_edge_se2(1000, 1001, [1, 0, 0], np.zeros(9))
    ctx.graph_set_fixed(1000)
    before = ctx.graph_get_vertex(1001).copy()
    assert ctx.graph_optimize_rc(3)[0] == 0                      # zero pivot -> g2o's "0 iterations"
    assert np.array_equal(ctx.graph_get_vertex(1001), before)
    with pytest.raises(Exception):
        ctx.graph_add_pose(1000, 0, 0, 0)                        # duplicate id
    with pytest.raises(Exception):
        ctx.graph_add_edge_se2_xy(1000, 77, [0, 0], np.eye(2).ravel())   # unknown landmark


def test_repeated_optimize_calls_continue_from_device_state(ctx, orc, synth):
    """The reference bursts optimizeGraph() once per remaining column (slam.cpp:625-633)."""
    g = small_graph(synth, 80)
    ctx.graph_load(g); G = orc.graph_from_soa(g)
    for _ in range(3):
        n, chi2 = ctx.graph_optimize(2); no, chi2o = G.optimize(2)
        assert n == no == 2 and np.allclose(chi2, chi2o, rtol=1e-8)
    pe, le = ctx.graph_get_estimates(); po, lo = G.estimates(g)
    assert _rel_close(pe, po) and _rel_close(le, lo)


def test_batched_replicas_match_oracle(ctx, orc, synth):
    g = small_graph(synth, 100)
    R = 6
    pe, le, ez, oz = synth.perturb_replicas(g, R, seed=18)
    ctx.graph_load(g)
    dpe, dle, chi2, done = ctx.graph_optimize_batch(pe, le, oz, ez, iters=10)
    assert np.all(done == 10)
    import copy
    for r in range(R):
        gr = copy.copy(g)
        gr.pose_est, gr.lm_est, gr.el_z, gr.eo_z = pe[r], le[r], ez[r], oz[r]
        G = orc.graph_from_soa(gr)
        no, chi2o = G.optimize(10)
        po, lo = G.estimates(gr)
        assert np.allclose(chi2[r], chi2o, rtol=1e-8), r
        assert _rel_close(dpe[r], po) and _rel_close(dle[r], lo), r


def test_optimize_c2_full_size(ctx, orc, synth):
    """BASELINE config 2 (10 laps, ~10k poses, 300 cones, single graph) at full size."""
    g = synth.c2_graph()
    ctx.graph_load(g)
    n, chi2 = ctx.graph_optimize(10)
    G = orc.graph_from_soa(g)
    no, chi2o = G.optimize(10)
    assert n == no == 10
    assert np.allclose(chi2, chi2o, rtol=1e-8)
    pe, le = ctx.graph_get_estimates(); po, lo = G.estimates(g)
    assert _rel_close(pe, po) and _rel_close(le, lo)
    # size-independent property: a converged GN step leaves chi2 stationary
    assert abs(chi2[-1] - chi2[-2]) <= 1e-9 * chi2[-1]


def test_device_map_update_and_pinned_mirror(ctx, orc, synth):
    """Slam::updateMap (slam.cpp:713-732) as a device kernel behind the optimise + the pinned host
    mirror of the cone map that sendCones / drawCones read.  The map built by the mapping-phase frames of a short
    drive, its graph optimised: every cone must take the estimate of landmark vertex j (device path), the mirror must
    equal the device map, and with NEWER host values (set_values after the optimise) the host estimates win."""
    trk = synth.ellipse_track()
    d = synth.simulate_drive(trk, 80, s_step=trk.length / 1000, seed=9)
    ctx.map_clear(); ctx.graph_clear()
    cci = lc = 0
    for fr, p in zip(d.frames, d.poses_noisy):
        r = ctx.assoc_map_frame(fr, p, synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD, cci, lc)
        cci, lc = r["cci"], r["loop_closing"]
    M = ctx.map_size()
    g = synth.graph_from_drive(d)
    assert len(g.lm_ids) <= M and np.array_equal(g.lm_ids, np.arange(len(g.lm_ids)))   # landmark id = map index
    ctx.graph_load(g)
    n, _ = ctx.graph_optimize(5)
    assert n == 5
    x0, y0, t0 = ctx.map_read()
    updated = ctx.map_update_from_graph()
    assert updated == len(g.lm_ids)
    pe, le = ctx.graph_get_estimates()
    x1, y1, t1 = ctx.map_read()
    L = len(g.lm_ids)
    assert np.array_equal(x1[:L], le[:, 0]) and np.array_equal(y1[:L], le[:, 1])      # bit for bit: a copy
    assert np.array_equal(x1[L:], x0[L:]) and np.array_equal(t1, t0)                  # cones without a vertex untouched
    mx, my, mt = ctx.map_mirror()
    assert np.array_equal(mx, x1) and np.array_equal(my, y1) and np.array_equal(mt, t1)
    # host values newer than the device's: they are what the map takes
    le2 = le + 0.25
    ctx.graph_set_values(pe, le2, None, None)
    assert ctx.map_update_from_graph() == L
    x2, y2, _ = ctx.map_read()
    assert np.array_equal(x2[:L], le2[:, 0]) and np.array_equal(y2[:L], le2[:, 1])
    mx, my, _ = ctx.map_mirror()
    assert np.array_equal(mx, x2) and np.array_equal(my, y2)
    # the mirror follows a map that grows without an update call
    ctx.map_append(np.array([1.0]), np.array([2.0]), np.array([3], dtype=np.int32))
    mx, my, mt = ctx.map_mirror()
    assert len(mx) == M + 1 and mx[-1] == 1.0 and my[-1] == 2.0 and mt[-1] == 3
    ctx.map_clear(); ctx.graph_clear()


def _pose_increments(pe):
    """prev^-1 * cur of consecutive poses (g2o SE2): the locally determined part of a pose chain."""
    d = pe[1:, :2] - pe[:-1, :2]
    c, s = np.cos(pe[:-1, 2]), np.sin(pe[:-1, 2])
    dth = pe[1:, 2] - pe[:-1, 2]
    dth = (dth + np.pi) % (2 * np.pi) - np.pi
    return np.stack([c * d[:, 0] + s * d[:, 1], -s * d[:, 0] + c * d[:, 1], dth], axis=1)


def test_corridor_graph_solve_matches_oracle(ctx, orc, synth):
    """BASELINE config 5's topology (poses every 0.3 m along a corridor, cone pairs every 3 m, each pose sees the cones
    1.5-12 m ahead) at 30,000 poses / 102k unknowns: a long, thin graph with a deep assembly tree (> 100 levels), solved
    by the same host analysis + front kernels as the track graphs.  The full 3.4M-unknown graph runs in bench.py
    (`c5.solve`, parity at 100k poses); the host analysis of the full size runs on the CPU under the stub runtime.

    An open 9 km chain held at one end is ill-conditioned along its bending modes (measured: two elimination orders of
    the same system move the far end by centimetres to metres while chi2 agrees to 1e-9), so absolute coordinates are
    compared through what the problem determines: chi2 per iteration, the chi2 THE ORACLE computes for the estimates
    of this path, and the locally determined quantities (consecutive-pose increments, landmarks in the frame of a
    pose that sees them) at the 1e-6 bar."""
    g = synth.c5_graph(n_poses=30_000, n_pairs=3_000)
    ctx.graph_load(g)
    n, chi2 = ctx.graph_optimize(6)
    st = ctx.graph_stats()
    assert st["n_levels"] > 50 and st["n"] == 3 * 30_000 + 2 * len(g.lm_ids) - 10
    G = orc.graph_from_soa(g)
    no, chi2o = G.optimize(6)
    assert n == no == 6
    assert np.allclose(chi2, chi2o, rtol=1e-4)              # the first steps travel along the weak modes too
    assert np.allclose(chi2[-2:], chi2o[-2:], rtol=1e-9)    # the same minimum
    pe, le = ctx.graph_get_estimates(); po, lo = G.estimates(g)
    import copy
    g2 = copy.copy(g)
    g2.pose_est, g2.lm_est = pe.copy(), le.copy()
    assert abs(orc.graph_from_soa(g2).chi2() - chi2o[-1]) <= 1e-9 * chi2o[-1]   # the oracle's own chi2 of OUR estimates
    assert np.max(np.abs(_pose_increments(pe) - _pose_increments(po))) <= 1e-6
    # every landmark in the frame of the first pose that sees it
    lm_index = {int(v): k for k, v in enumerate(g.lm_ids)}
    pose_index = {int(v): k for k, v in enumerate(g.pose_ids)}
    first_pose = {}
    for p, l in zip(g.el_pose, g.el_lm):
        first_pose.setdefault(int(l), int(p))
    ls = np.array([lm_index[l] for l in first_pose]); ps = np.array([pose_index[p] for p in first_pose.values()])

    def local(pev, lev):
        d = lev[ls] - pev[ps, :2]
        c, s = np.cos(pev[ps, 2]), np.sin(pev[ps, 2])
        return np.stack([c * d[:, 0] + s * d[:, 1], -s * d[:, 0] + c * d[:, 1]], axis=1)
    assert np.max(np.abs(local(pe, le) - local(po, lo))) <= 1e-6


def test_fronts_beyond_shared_memory(ctx, orc, synth, monkeypatch):
    """Fronts that do not fit an SM's shared memory (> 163 rows) are factorised in a global-memory slab
    (factor2_kernel<false>, backward_kernel<false>).  The default ordering never produces them on the
    trackdrive graphs, so the ordering is forced to leaf-level dissection (landmarks eliminated early,
    one dense clique per landmark): 6 laps x 300 poses -> fronts of up to ~194 rows."""
    monkeypatch.setenv("SLAM_B200_ND_LEAF", "8")
    g = synth.c2_graph(n_laps=6, poses_per_lap=300)
    ctx.graph_load(g)
    n, chi2 = ctx.graph_optimize(10)
    st = ctx.graph_stats()
    assert st["max_front"] > 163 and st["nFbig"] > 0, st   # the path under test was really taken
    G = orc.graph_from_soa(g)
    no, chi2o = G.optimize(10)
    assert n == no == 10
    assert np.allclose(chi2, chi2o, rtol=1e-8)
    pe, le = ctx.graph_get_estimates(); po, lo = G.estimates(g)
    assert _rel_close(pe, po) and _rel_close(le, lo)
    monkeypatch.delenv("SLAM_B200_ND_LEAF")
    ctx.graph_load(small_graph(synth, 40))  # leave the context with a default-ordered structure
    ctx.graph_optimize(1)


@pytest.mark.parametrize("variant", ["1"])
def test_first_generation_front_kernel_still_agrees(ctx, orc, synth, variant):
    """SLAM_B200_FACTOR_VARIANT=1 (kept for A/B measurements) is read once per process, so it runs in a
    child process: same graph, same tolerance."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from conftest import load_pkg, small_graph\n"
        "from oracle import oracle\n"
        "pkg = load_pkg(); orc = oracle.load('best'); g = small_graph(pkg.synth, 300)\n"
        "ctx = pkg.Context(0); ctx.graph_load(g); n, chi2 = ctx.graph_optimize(10)\n"
        "G = orc.graph_from_soa(g); no, chi2o = G.optimize(10)\n"
        "pe, le = ctx.graph_get_estimates(); po, lo = G.estimates(g)\n"
        "assert n == no == 10 and np.allclose(chi2, chi2o, rtol=1e-8)\n"
        "assert np.max(np.abs(pe - po)) <= 1e-6 * max(1.0, np.max(np.abs(po)))\n"
        "assert np.max(np.abs(le - lo)) <= 1e-6 * max(1.0, np.max(np.abs(lo)))\n"
        "print('ok')\n"
    ) % (os.path.dirname(__file__), os.path.dirname(os.path.dirname(__file__)))
    env = dict(os.environ, SLAM_B200_FACTOR_VARIANT=variant)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


_VARIANT_CODE = (
    "import sys, hashlib, json, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
    "from conftest import load_pkg\n"
    "pkg = load_pkg(); g = pkg.synth.c2_graph(n_laps=4, poses_per_lap=400)\n"
    "ctx = pkg.Context(0); ctx.graph_load(g); n, chi2 = ctx.graph_optimize(10)\n"
    "st = ctx.graph_stats(); assert st['max_front'] > 64 and n == 10, (st, n)\n"
    "pe, le = ctx.graph_get_estimates()\n"
    "np.save(sys.argv[1], np.concatenate([pe.ravel(), le.ravel(), np.asarray(chi2)]))\n"
    "print('ok', int(ctx.launch_count()))\n"
)


def _run_variant(tmp_path, name, env_extra):
    import subprocess
    import sys
    out = str(tmp_path / (name + ".npy"))
    code = _VARIANT_CODE % (os.path.dirname(__file__), os.path.dirname(os.path.dirname(__file__)))
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", code, out], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
    return np.load(out), int(r.stdout.split()[-1])


def test_launch_structure_variants_are_bit_identical(tmp_path):
    """The single-graph front kernels are chained by programmatic dependent launch (prologue of a level under the level
    before it), mixed levels go out as one launch and the roots solve their backward system inside the factor kernel.
    Neither the chaining nor the fused roots change a single operation or its order: with them switched off (environment
    read once per process, hence child processes) the estimates and chi2 of a 4-lap graph (fronts beyond 64 rows, ten
    levels) after ten iterations are the SAME BITS.  A prologue that read something the previous launch was still writing, or
    a stale line of it, would show up here."""
    base, n_base = _run_variant(tmp_path, "default", {})
    for name, env in (("no_pdl", {"SLAM_B200_NO_PDL": "1"}), ("no_root_fuse", {"SLAM_B200_NO_ROOT_FUSE": "1"}),
                      ("no_graph", {"SLAM_B200_NO_CUDA_GRAPH": "1"})):
        other, n_other = _run_variant(tmp_path, name, env)
        assert base.shape == other.shape and np.array_equal(base, other), (name, float(np.max(np.abs(base - other))))
        if name == "no_root_fuse":
            assert n_other > n_base  # the root level's backward launch is back
    again, _ = _run_variant(tmp_path, "again", {})
    assert np.array_equal(base, again)
    # a mixed level as two launches sends its small fronts through the warp-per-front backward kernel, whose sums run
    # in another order: same answer to rounding, not the same bits
    split, n_split = _run_variant(tmp_path, "no_merge", {"SLAM_B200_NO_LEVEL_MERGE": "1"})
    assert n_split > n_base
    assert np.max(np.abs(base - split)) <= 1e-11 * max(1.0, float(np.max(np.abs(base))))


def test_tensor_pipe_and_scalar_trailing_update_agree(tmp_path):
    """The trailing update of a front runs as 8 x 8 tiles of mma.sync.m8n8k4.f64 straight from shared memory (the next
    triangle factorised by shuffles in the accumulator layout, seeded reciprocals); SLAM_B200_UPDATE_MMA=0 is the scalar
    update with the redundant triangle and correctly rounded reciprocals.  Different rounding, same answer: estimates
    within 1e-9 relative, chi2 within 1e-10 relative."""
    a, _ = _run_variant(tmp_path, "mma", {})
    b, _ = _run_variant(tmp_path, "scalar", {"SLAM_B200_UPDATE_MMA": "0"})
    assert np.max(np.abs(a[:-10] - b[:-10])) <= 1e-9 * max(1.0, float(np.max(np.abs(b[:-10]))))
    assert np.allclose(a[-10:], b[-10:], rtol=1e-10)


def test_front_timeline_is_consistent(tmp_path):
    """SLAM_B200_TIMELINE: every front of the last iteration carries ordered stamps (entry <= wait passed <= ... <= end),
    a parent's dependency wait ends after the end of each of its children, and the fused roots have no backward stamps."""
    import subprocess
    import sys
    code = (
        "import sys, ctypes as C, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from conftest import load_pkg\n"
        "pkg = load_pkg(); ctx = pkg.Context(0); ctx.graph_load(pkg.synth.graph_from_drive(pkg.synth.trackdrive(1)))\n"
        "ctx.graph_prepare(); ctx.graph_iterate_async(3); ctx.sync(); L = ctx.L\n"
        "L.slam_b200_debug_timeline.restype = C.c_long; L.slam_b200_debug_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_long]\n"
        "n = L.slam_b200_debug_timeline(ctx.h, None, 0); assert n > 0 and n %% 12 == 0, n\n"
        "buf = (C.c_longlong * n)(); assert L.slam_b200_debug_timeline(ctx.h, buf, n) == n\n"
        "t = np.array(buf, dtype=np.int64).reshape(-1, 12)\n"
        "L.slam_b200_graph_export_symbolic.restype = C.c_long\n"
        "L.slam_b200_graph_export_symbolic.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_long]\n"
        "m = L.slam_b200_graph_export_symbolic(ctx.h, 7, None, 0); par = np.zeros(m, dtype=np.int32)\n"
        "L.slam_b200_graph_export_symbolic(ctx.h, 7, par.ctypes.data_as(C.POINTER(C.c_int32)), m)\n"
        "assert len(par) == len(t)\n"
        "assert np.all(t[:, 0] > 0) and np.all(np.diff(t[:, :6], axis=1) >= 0)\n"
        "for f, p in enumerate(par):\n"
        "    if p >= 0: assert t[p, 1] >= t[f, 5], (f, p)\n"
        "    else: assert t[f, 6] == 0 and t[f, 8] == 0, f\n"
        "    if p >= 0: assert 0 < t[f, 6] <= t[f, 7] <= t[f, 8] and t[f, 7] >= max(t[p, 5], t[p, 8]), f\n"
        "print('ok')\n"
    ) % (os.path.dirname(__file__), os.path.dirname(os.path.dirname(__file__)))
    env = dict(os.environ, SLAM_B200_TIMELINE="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])


def test_register_resident_large_front_kernel_agrees(ctx, orc, synth):
    """SLAM_B200_FACTOR_VARIANT=3: the large fronts (more than 64 rows) of a single graph by factor3_kernel -- the tile
    triangle in the registers of 16 warps, DMMA updates (csrc/factor3.cuh; measured no faster than factor2_kernel and
    therefore off by default).  Read once per process, so it runs in a child process: a 4-lap graph whose root fronts
    have ~100 rows, same tolerances as the default path."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from conftest import load_pkg\n"
        "from oracle import oracle\n"
        "pkg = load_pkg(); orc = oracle.load('best'); g = pkg.synth.c2_graph(n_laps=4, poses_per_lap=400)\n"
        "ctx = pkg.Context(0); ctx.graph_load(g); n, chi2 = ctx.graph_optimize(10)\n"
        "st = ctx.graph_stats(); assert st['max_front'] > 64, st\n"
        "G = orc.graph_from_soa(g); no, chi2o = G.optimize(10)\n"
        "pe, le = ctx.graph_get_estimates(); po, lo = G.estimates(g)\n"
        "assert n == no == 10 and np.allclose(chi2, chi2o, rtol=1e-8)\n"
        "assert np.max(np.abs(pe - po)) <= 1e-6 * max(1.0, np.max(np.abs(po)))\n"
        "assert np.max(np.abs(le - lo)) <= 1e-6 * max(1.0, np.max(np.abs(lo)))\n"
        "print('ok')\n"
    ) % (os.path.dirname(__file__), os.path.dirname(os.path.dirname(__file__)))
    env = dict(os.environ, SLAM_B200_FACTOR_VARIANT="3")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_sharded_assembly_partials_sum_to_full(ctx, pkg, synth):
    """Edge-partitioned assembly (config 5): assembling pose ranges separately gives pose blocks and
    off-diagonal blocks owned by exactly one shard and landmark partial sums that add up to the full
    assembly (the part ranks all-reduce)."""
    import importlib
    import torch
    par = importlib.import_module(pkg.__name__ + ".parallel")
    g = small_graph(synth, 200)
    ctx.graph_load(g)
    ctx.graph_prepare_assembly_only()
    P, L = len(g.pose_ids), len(g.lm_ids)
    ptr, nV = ctx.graph_system_dev(1)
    V = torch.as_tensor(par.DeviceArray(ptr, nV), device="cuda")
    V.zero_(); torch.cuda.synchronize()          # slots of fixed vertices are never written
    ctx.graph_assemble_async(0, P); ctx.sync()
    full = V.clone()
    parts = []
    for r in range(3):
        lo, hi = par.shard_range(P, r, 3)
        V.zero_(); torch.cuda.synchronize()
        ctx.graph_assemble_async(lo, hi); ctx.sync()
        parts.append(V.clone())
    tot = parts[0] + parts[1] + parts[2]
    assert torch.allclose(tot[:6 * L], full[:6 * L], rtol=1e-12, atol=1e-14)        # landmark part: sums
    # everything else has exactly one owner shard (other shards leave zeros); the warp-level
    # summation tree depends on where a shard starts, so last-bit differences are allowed
    assert torch.allclose(tot[6 * L:], full[6 * L:], rtol=1e-13, atol=1e-15)
    owners = sum((p[6 * L:] != 0).to(torch.int32) for p in parts)
    assert int(owners.max()) <= 1
    lm, n = ctx.graph_system_dev(0)
    assert n == 6 * L and lm == ptr


def test_two_contexts_on_two_host_threads():
    """Two contexts used from two host threads at the same time (own process: tests/gpu_case_two_contexts.py) give
    bit for bit what each gives alone -- the process-wide host pool of the symbolic phase, the structure upload
    list and the kernel attributes are the shared pieces this guards."""
    import os
    import subprocess
    import sys
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "gpu_case_two_contexts.py")], capture_output=True, text=True, timeout=420)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def _run_case(env_extra, prefix=()):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SLAM_B200_SANITIZER_SMALL="1", **env_extra)
    return subprocess.run(list(prefix) + [sys.executable, os.path.join(root, "tests", "gpu_case_sanitizer.py")],
                          capture_output=True, text=True, timeout=420, env=env, cwd=root)


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_kernels_under_compute_sanitizer(tool):
    """compute-sanitizer memcheck / racecheck over a small end-to-end case (tests/gpu_case_sanitizer.py:
    association frames, grid and brute-force bulk association, one graph optimisation through the CTA and warp
    front kernels, a replica batch) in its own process: no invalid access, no shared-memory hazard.
    On the pool this repo is measured on the tool is CLOSED by the operators (it exits 86 with a notice;
    profiles/r02_compute_sanitizer_closed.log): the test then skips and test_kernels_under_guard_bands /
    test_case_is_bit_reproducible below are the checks that run."""
    import shutil
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cs = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(cs):
        pytest.skip("compute-sanitizer not installed")
    r = _run_case({}, prefix=(cs, "--tool", tool, "--error-exitcode", "3"))
    out = r.stdout + r.stderr
    if "closed on this pool" in out:
        pytest.skip("compute-sanitizer is closed on this pool by its operators: " + out.strip()[:160])
    # --error-exitcode makes any finding a non-zero exit; the banner proves the tool really ran
    assert r.returncode == 0 and "COMPUTE-SANITIZER" in out and "batch" in out, out[-3000:]


def _verdict(r):
    import json
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_kernels_under_guard_bands():
    """The library's own memory check (SLAM_B200_GUARD=1, csrc/ctx.h): every device array sits between two 4 KiB
    bands of 0xFF and starts out as 0xFF itself (fp64 NaN / int32 -1).  The case compares every result with the
    CPU oracle, so an out-of-bounds or uninitialised read poisons a comparison, and at the end no band byte may
    have changed (no out-of-bounds write)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    v = _verdict(_run_case({"SLAM_B200_GUARD": "1"}))
    assert v["guard_mode"] and v["arrays_checked"] > 40 and v["guard_bytes_changed"] == 0, v


def test_case_is_bit_reproducible():
    """No atomics on any value path and fixed summation orders: two runs of the case (one of them with the guard
    bands, i.e. with a different memory layout and NaN-filled fresh allocations) must give bit-identical results --
    a shared-memory or global-memory race is the only thing that could make them differ."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    a = _verdict(_run_case({}))
    b = _verdict(_run_case({}))
    c = _verdict(_run_case({"SLAM_B200_GUARD": "1"}))
    assert a["hash"] == b["hash"] == c["hash"], (a, b, c)


def test_sparse_vertex_ids_and_warmup_do_not_change_results(pkg, synth):
    """The id index is a flat table while the ids are dense and a hash map when they are sparse (csrc/ctx.h: IdIndex);
    slam_b200_warmup optimises a synthetic ring first and must leave the context empty.  The same graph with the
    reference's numbering, with ids scattered over 2^30 (bulk load and incremental API), and after a warm-up gives
    bit-identical estimates: none of this touches the arithmetic."""
    import copy
    g = small_graph(synth, 120)
    ctx = pkg.Context(0)
    ctx.graph_load(g)
    n0, chi0 = ctx.graph_optimize(5)
    pe0, le0 = ctx.graph_get_estimates()
    ctx.close()
    rng = np.random.default_rng(3)
    # order-preserving sparse renumbering (the Hessian order is by ascending id)
    all_ids = np.sort(np.concatenate([g.lm_ids, g.pose_ids]))
    new = np.sort(rng.choice(np.arange(1, 1 << 30), size=len(all_ids), replace=False)).astype(np.int32)
    remap = dict(zip(all_ids.tolist(), new.tolist()))
    m = np.vectorize(remap.get)
    g2 = copy.copy(g)
    g2.lm_ids, g2.pose_ids = m(g.lm_ids).astype(np.int32), m(g.pose_ids).astype(np.int32)
    g2.eo_from, g2.eo_to = m(g.eo_from).astype(np.int32), m(g.eo_to).astype(np.int32)
    g2.el_pose, g2.el_lm = m(g.el_pose).astype(np.int32), m(g.el_lm).astype(np.int32)
    g2.fixed_ids = m(g.fixed_ids).astype(np.int32)
    ctx = pkg.Context(0)
    ctx.warmup(256, 80)
    assert ctx.graph_num_poses() == 0 and ctx.graph_num_landmarks() == 0 and ctx.map_size() == 0
    ctx.graph_load(g2)
    n1, chi1 = ctx.graph_optimize(5)
    pe1, le1 = ctx.graph_get_estimates()
    assert n0 == n1 == 5 and np.array_equal(chi0, chi1) and np.array_equal(pe0, pe1) and np.array_equal(le0, le1)
    assert np.array_equal(ctx.graph_get_vertex(int(g2.pose_ids[7])), pe1[7])
    # incremental API with the sparse ids, poses first (the id range grows downwards and upwards)
    ctx.graph_clear()
    for i, vid in enumerate(g2.pose_ids):
        ctx.graph_add_pose(int(vid), *g2.pose_est[i])
    for i, vid in enumerate(g2.lm_ids):
        ctx.graph_add_landmark(int(vid), *g2.lm_est[i])
    for e in range(len(g2.eo_from)):
        ctx.graph_add_edge_se2(int(g2.eo_from[e]), int(g2.eo_to[e]), g2.eo_z[e], g2.eo_info[e])
    for e in range(len(g2.el_pose)):
        ctx.graph_add_edge_se2_xy(int(g2.el_pose[e]), int(g2.el_lm[e]), g2.el_z[e], g2.el_info[e])
    for vid in g2.fixed_ids:
        ctx.graph_set_fixed(int(vid), True)
    n2, chi2 = ctx.graph_optimize(5)
    pe2, le2 = ctx.graph_get_estimates()
    assert n2 == 5 and np.allclose(chi2, chi0, rtol=1e-12) and np.allclose(pe2, pe0, rtol=0, atol=1e-12)
    ctx.close()

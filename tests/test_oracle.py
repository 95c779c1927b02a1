"""The oracle itself (CPU only): known answers, internal cross-checks (finite differences, scipy,
port-vs-Eigen), and the committed golden fixtures.  PARITY UNPINNED by the reference's own tests
(it has none for this path); these are the pins we do have."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import small_graph

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_known_answers(orc):
    ka = json.load(open(os.path.join(GOLD, "known_answers.json")))
    assert repr(orc.lib.orc_pi_ref()) == ka["pi_ref"]
    for case in ka["spherical2cartesian"]:
        out = orc.spherical2cartesian(*case["in"])
        for got, want in zip(out, case["out"]):
            if want is None:
                assert np.isnan(got)
            else:
                assert got == pytest.approx(want, rel=0, abs=1e-15)
    for case in ka["cone_to_global"]:
        out = orc.cone_to_global(case["pose"], case["obs"])
        assert out[0] == pytest.approx(case["out"][0], abs=2e-15)
        assert out[1] == pytest.approx(case["out"][1], abs=2e-15)


def test_normalize_theta(orc):
    f = orc.lib.orc_normalize_theta
    assert f(0.5) == 0.5
    assert f(np.pi) == pytest.approx(-np.pi)
    assert f(-np.pi) == -np.pi
    assert f(7.0) == pytest.approx(7.0 - 2 * np.pi)
    assert f(-7.0) == pytest.approx(-7.0 + 2 * np.pi)


def test_jacobians_finite_difference(orc, synth):
    g = small_graph(synth, 30)
    G = orc.graph_from_soa(g)
    ne = G.L.orc_graph_num_edges(G.h)
    ids, est0 = G.get_all()
    checked = 0
    for e in list(range(0, ne, max(1, ne // 25)))[:25]:
        D, err, Ji, Jj = G.edge_linearization(e)
        # perturb every vertex coordinate, find the two that move this edge's error
        for v, vid in enumerate(ids):
            dim = 3 if vid >= 1000 else 2
            for k in range(dim):
                h = 1e-6
                x = est0[v].copy(); x[k] += h
                G.L.orc_graph_set_vertex(G.h, int(vid), x.ctypes.data_as(orc.lib.orc_graph_set_vertex.argtypes[2]) if False else x.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double)))
                _, e1, _, _ = G.edge_linearization(e)
                x[k] -= 2 * h
                G.L.orc_graph_set_vertex(G.h, int(vid), x.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double)))
                _, e2, _, _ = G.edge_linearization(e)
                x[k] += h
                G.L.orc_graph_set_vertex(G.h, int(vid), x.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_double)))
                fd = (e1 - e2)[:D] / (2 * h)
                if np.any(np.abs(fd) > 1e-9):
                    # this vertex belongs to the edge: its column must match Ji or Jj
                    ok = np.allclose(fd, Ji[:D, k], atol=1e-6) or np.allclose(fd, Jj[:D, k], atol=1e-6)
                    assert ok, (e, vid, k, fd, Ji, Jj)
                    checked += 1
    assert checked > 50


def test_linear_solve_against_scipy(orc, synth):
    g = small_graph(synth, 100)
    G = orc.graph_from_soa(g)
    s = G.build_system()
    n = s["n"]
    U = sp.csc_matrix((s["Ax"], s["Ai"], s["Ap"]), shape=(n, n))
    H = U + sp.triu(U, 1).T
    dx = spla.spsolve(sp.csc_matrix(H), s["b"])
    before = G.estimates(g)
    G.optimize(1)
    after = G.estimates(g)
    # one GN step moves every free vertex by the scipy solution (additive update)
    hidx = s["hidx"]
    ids, _ = G.get_all()
    pos = {int(i): k for k, i in enumerate(ids)}
    for k, vid in enumerate(g.pose_ids):
        h = hidx[pos[int(vid)]]
        d = after[0][k] - before[0][k]
        if h >= 0:
            assert np.allclose(d[:2], dx[h:h + 2], rtol=1e-7, atol=1e-10)
        else:
            assert np.all(d == 0)
    for k, vid in enumerate(g.lm_ids):
        h = hidx[pos[int(vid)]]
        d = after[1][k] - before[1][k]
        if h >= 0:
            assert np.allclose(d, dx[h:h + 2], rtol=1e-7, atol=1e-10)


def test_port_vs_reference_eigen(orc_port, synth):
    from oracle import oracle
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ref = oracle.load("reference")
    g = small_graph(synth, 200)
    A = orc_port.graph_from_soa(g); B = ref.graph_from_soa(g)
    na, ca = A.optimize(10); nb, cb = B.optimize(10)
    assert na == nb == 10
    assert np.allclose(ca, cb, rtol=1e-10)
    pa, la = A.estimates(g); pb, lb = B.estimates(g)
    assert np.allclose(pa, pb, rtol=1e-9, atol=1e-10) and np.allclose(la, lb, rtol=1e-9, atol=1e-10)


def test_noise_free_graph_converges_to_truth(orc, synth):
    trk = synth.ellipse_track()
    d = synth.simulate_drive(trk, 150, s_step=trk.length / 1000, seed=3, sigma_xy=0, sigma_th=0, sigma_r=0, sigma_az=0)
    g = synth.graph_from_drive(d)
    # perturb the free poses a little; measurements are exact (up to float32 rounding of az/range)
    rng = np.random.default_rng(0)
    g.pose_est[2:] += rng.normal(size=g.pose_est[2:].shape) * np.array([0.05, 0.05, 0.005])
    G = orc.graph_from_soa(g)
    n, chi2 = G.optimize(10)
    assert n == 10 and chi2[-1] < 1e-6
    pe, _ = G.estimates(g)
    assert np.allclose(pe[:, :2], d.poses_true[:, :2], atol=2e-3)


def test_gauge_and_return_codes(orc):
    G = orc.graph()
    assert G.optimize(3)[0] == -1                    # nothing to optimise
    G.add_pose(1000, 0, 0, 0); G.add_pose(1001, 1, 0, 0)
    G.add_edge_se2(1000, 1001, [1, 0, 0], np.eye(3).ravel() * 5)
    G.set_fixed(1000); G.set_fixed(1001)
    assert G.optimize(3)[0] == -1                    # all edges inactive
    G.set_fixed(1001, False)
    G.add_pose(1002, 2, 0, 0)                        # a free vertex without edges stays out
    assert G.optimize(3)[0] == 3
    # a zero information matrix gives a zero pivot -> g2o returns 0
    H = orc.graph()
    H.add_pose(1000, 0, 0, 0); H.add_pose(1001, 1, 0, 0)
    H.add_edge_se2(1000, 1001, [1, 0, 0], np.zeros(9))
    H.set_fixed(1000)
    assert H.optimize(3)[0] == 0


def test_c1_replay_golden(orc, synth, c1_drive):
    """C1: the trackdrive loop replayed through the restated Slam back half; outputs pinned in
    tests/golden/c1_replay.npz (generated by tests/golden/make_golden.py from this same oracle)."""
    gold = np.load(os.path.join(GOLD, "c1_replay.npz"))
    s = orc.slam(synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    idx_all, st_all, rcs = [], [], []
    for fr, p in zip(c1_drive.frames, c1_drive.poses_noisy):
        rc, idx, st = s.perform(fr, p)
        idx_all.append(idx); st_all.append(st); rcs.append(rc)
    assert np.array_equal(np.concatenate(idx_all), gold["idx"])
    assert np.array_equal(np.concatenate(st_all), gold["status"])
    assert np.array_equal(np.array(rcs), gold["rc"])
    mx, my, mt = s.map()
    assert np.allclose(mx, gold["map_x"], rtol=1e-9) and np.allclose(my, gold["map_y"], rtol=1e-9)
    assert np.array_equal(mt, gold["map_type"])
    assert np.allclose(s.chi2_log(), gold["chi2"], rtol=1e-8)
    st = s.state()
    assert st["loop_closing_complete"] == 1 and st["current_cone_index"] == int(gold["cci"])

"""Host symbolic phase (nested dissection + assembly tree): structural invariants and, through a
numpy emulation of the front schedule, that it really factorises and solves H x = b."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import small_graph, skip_if_sanitizer_runtime_unusable
import mf_emul


def _check_structure(sym, dims):
    n = int(np.sum(dims))
    assert sym.n == n and sym.nb == len(dims)
    assert sorted(sym.pos.tolist()) == list(range(len(dims)))
    # pivots tile the solver index space; children precede parents; levels are contiguous
    order = np.argsort(sym.piv0, kind="stable")
    assert np.array_equal(order, np.arange(sym.nf))
    assert sym.piv0[0] == 0 and np.all(sym.piv0[1:] == np.cumsum(sym.npiv)[:-1]) and np.sum(sym.npiv) == n
    for f in range(sym.nf):
        if sym.parent[f] >= 0:
            assert sym.parent[f] > f
        rows = sym.upd_rows[sym.rows_ptr[f]:sym.rows_ptr[f + 1]]
        assert np.all(np.diff(rows) > 0)
        if len(rows):
            assert rows[0] >= sym.piv0[f] + sym.npiv[f]
            assert sym.parent[f] >= 0
    lvl = np.zeros(sym.nf, dtype=int)
    for l in range(sym.nlevels):
        lvl[sym.level_ptr[l]:sym.level_ptr[l + 1]] = l
    for f in range(sym.nf):
        if sym.parent[f] >= 0:
            assert lvl[sym.parent[f]] > lvl[f]


@pytest.mark.parametrize("amalg", [None, "14,1.5,0.9,2.0,4.0,150"])
@pytest.mark.parametrize("leaf", [1, 8, 64, 100000])
def test_symbolic_solves_small_graph(pkg, synth, orc, leaf, amalg, monkeypatch):
    """amalg: the optional latency-driven amalgamation (children merged into their parents across
    siblings, symbolic.cpp) must give a structure that still factorises and solves exactly."""
    if amalg:
        monkeypatch.setenv("SLAM_B200_AMALG", amalg)
    g = small_graph(synth, 150)
    ids, dims, pa, pb = mf_emul.block_pattern(g)
    sym = pkg.SymbolicAnalysis(dims, pa, pb, leaf_size=leaf)
    _check_structure(sym, dims)
    G = orc.graph_from_soa(g)
    sysm = G.build_system()
    n = sysm["n"]
    assert n == sym.n
    U = sp.csc_matrix((sysm["Ax"], sysm["Ai"], sysm["Ap"]), shape=(n, n))
    H = (U + sp.triu(U, 1).T).toarray()
    hv = mf_emul.hvals_from_dense(H, dims, pa, pb, sym)
    perm = mf_emul.solver_perm(sym, dims)
    x_solver = mf_emul.factor_solve(sym, hv, sysm["b"][perm])
    x = np.zeros(n)
    x[perm] = x_solver
    x_ref = spla.spsolve(sp.csc_matrix(H), sysm["b"])
    # normwise: cond(H) ~ 2e6 on this graph, so entries near zero carry ~1e-11 of absolute error in
    # either solver; the elimination order (and with it the rounding) changes with the amalgamation
    assert np.max(np.abs(x - x_ref)) <= 1e-9 * np.max(np.abs(x_ref))


def test_symbolic_c1_quality(pkg, synth, c1_graph):
    """The ordering must stay sparse on the trackdrive topology (never a dense Schur)."""
    ids, dims, pa, pb = mf_emul.block_pattern(c1_graph)
    sym = pkg.SymbolicAnalysis(dims, pa, pb, leaf_size=1024)
    _check_structure(sym, dims)
    n = sym.n
    assert sym.nnzL < 40 * n          # AMD on the reference solver gives ~20 n on this pattern
    assert sym.nlevels < 40           # short assembly tree = few dependent kernel levels
    assert sym.max_front < 400


def test_symbolic_random_sparse(pkg):
    """Generic patterns (disconnected pieces, cliques, isolated vertices) are handled."""
    rng = np.random.default_rng(3)
    nb = 200
    dims = rng.integers(2, 4, nb).astype(np.int32)
    pairs = set()
    for _ in range(500):
        a, b = rng.integers(0, nb - 20, 2)   # leaves the last 20 vertices isolated
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    for a in range(60, 70):                  # a clique
        for b in range(a + 1, 70):
            pairs.add((a, b))
    pairs = sorted(pairs)
    pa = np.array([p[0] for p in pairs], dtype=np.int32); pb = np.array([p[1] for p in pairs], dtype=np.int32)
    sym = pkg.SymbolicAnalysis(dims, pa, pb, leaf_size=4)
    _check_structure(sym, dims)
    n = int(dims.sum())
    off = np.concatenate([[0], np.cumsum(dims)])
    H = np.zeros((n, n))
    for a, b in pairs:
        blk = rng.normal(size=(dims[a], dims[b])) * 0.1
        H[off[a]:off[a + 1], off[b]:off[b + 1]] = blk
        H[off[b]:off[b + 1], off[a]:off[a + 1]] = blk.T
    H += np.eye(n) * (np.abs(H).sum(axis=1).max() + 1.0)
    b = rng.normal(size=n)
    hv = mf_emul.hvals_from_dense(H, dims, pa, pb, sym)
    perm = mf_emul.solver_perm(sym, dims)
    x = np.zeros(n)
    x[perm] = mf_emul.factor_solve(sym, hv, b[perm])
    assert np.allclose(H @ x, b, rtol=1e-9, atol=1e-10)


def test_symbolic_is_independent_of_the_host_thread_count():
    """Nested dissection below depth 1 and the region-wise minimum degree run on a pool of host threads;
    the elimination order, the assembly tree and every index array must not depend on how many."""
    import hashlib
    import os
    import subprocess
    import sys
    code = r'''
import sys, hashlib, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import load_pkg
import mf_emul
pkg = load_pkg()
g = pkg.synth.graph_from_drive(pkg.synth.trackdrive(5))
ids, dims, pa, pb = mf_emul.block_pattern(g)
S = pkg.capi.SymbolicAnalysis(dims, pa, pb, 512)
h = hashlib.sha1()
for k in ("pos", "boff", "level_ptr", "npiv", "nupd", "parent", "rel", "asm"):
    h.update(np.ascontiguousarray(getattr(S, k)).tobytes())
print("HASH", h.hexdigest(), int(S.nnzL))
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    seen = set()
    # ... nor on which minimum-degree implementation ran (dense bit matrix vs sorted adjacency lists)
    for threads, sparse in (("1", None), ("2", None), ("7", None), ("3", "1")):
        env = dict(os.environ, SLAM_B200_SYM_THREADS=threads)
        env.pop("SLAM_B200_MD_SPARSE", None)
        if sparse:
            env["SLAM_B200_MD_SPARSE"] = sparse
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        seen.add([ln for ln in out.stdout.splitlines() if ln.startswith("HASH")][0])
    assert len(seen) == 1, seen


@pytest.mark.parametrize("san", ["address,undefined", "thread"])
def test_symbolic_phase_under_sanitizers(pkg, tmp_path, san):
    """csrc/symbolic.cpp under AddressSanitizer + UBSan and under ThreadSanitizer (the phase runs on a pool of host
    threads): SLAM-shaped and adversarial patterns, two repetitions whose results must be identical."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    csrc = os.path.join(os.path.dirname(here), pkg.__name__, "csrc")
    exe = str(tmp_path / "sanitize_symbolic")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=" + san, "-fno-sanitize-recover=all", "-I" + csrc,
           os.path.join(here, "sanitize_symbolic_driver.cpp"), os.path.join(csrc, "symbolic.cpp"), "-o", exe, "-lpthread"]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0:
        pytest.skip("sanitizer runtime not available: " + b.stderr[-300:])
    env = dict(os.environ, SLAM_B200_SYM_THREADS="8", TSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=240, env=env)
    skip_if_sanitizer_runtime_unusable(r.stdout + r.stderr)
    assert r.returncode == 0 and r.stdout.strip() == "ok 0", r.stdout[-500:] + r.stderr[-3000:]


@pytest.mark.parametrize("n_poses,leaf", [(150, 1024), (60, 8), (400, 1024), (1000, 1024), (1000, 64)])
def test_tile_plan_solves(pkg, synth, orc, n_poses, leaf):
    """Host plan of the tiled batched factorisation (csrc/tileplan.cpp): local layout with identity padding and the
    rhs as an extra row, tile-major storage, assembly items from V and from the children's stored fronts -- run through
    the numpy emulation of what one warp of factor_tile_kernel / backward_tile_kernel does, it must solve H x = b."""
    g = small_graph(synth, n_poses)
    ids, dims, pa, pb = mf_emul.block_pattern(g)
    sym = pkg.SymbolicAnalysis(dims, pa, pb, leaf_size=leaf)
    if not sym.tile_ok:
        pytest.skip("a front of this ordering exceeds 96 local rows")
    G = orc.graph_from_soa(g)
    sysm = G.build_system()
    n = sysm["n"]
    U = sp.csc_matrix((sysm["Ax"], sysm["Ai"], sysm["Ap"]), shape=(n, n))
    H = (U + sp.triu(U, 1).T).toarray()
    hv = mf_emul.hvals_from_dense(H, dims, pa, pb, sym)
    assert sym.tile_rhs_base == len(hv)
    perm = mf_emul.solver_perm(sym, dims)
    x_solver = mf_emul.tile_factor_solve(sym, hv, sysm["b"][perm])
    x = np.zeros(n)
    x[perm] = x_solver
    x_ref = spla.spsolve(sp.csc_matrix(H), sysm["b"])
    assert np.max(np.abs(x - x_ref)) <= 1e-9 * np.max(np.abs(x_ref))
    # the same answer as the general front schedule
    x_gen = np.zeros(n)
    x_gen[perm] = mf_emul.factor_solve(sym, hv, sysm["b"][perm])
    assert np.max(np.abs(x - x_gen)) <= 1e-9 * np.max(np.abs(x_ref))


def test_tile_plan_c1(pkg, synth, c1_graph):
    """Config 3's topology (the 1-lap graph, batch ordering) must be a tiled topology, with fronts of <= 64 local rows."""
    ids, dims, pa, pb = mf_emul.block_pattern(c1_graph)
    sym = pkg.SymbolicAnalysis(dims, pa, pb, leaf_size=1024)
    assert sym.tile_ok and sym.tile_max_T <= 8
    assert np.all(sym.tile_item_nv % 32 == 0) and np.all(np.diff(sym.tile_item_ptr) % 32 == 0)

"""numpy emulation of the multifrontal LDL^T schedule the CUDA kernels execute (solver.cu), driven
by the arrays of the host symbolic analysis.  Test infrastructure: it validates the symbolic phase
on a CPU-only box; it is not part of the product."""
import numpy as np


def block_pattern(g):
    """Free vertices in g2o order (ascending id), dims, distinct pairs (a<b) of a GraphSoA."""
    fixed = set(int(v) for v in g.fixed_ids)
    ids = [(int(v), 2) for v in g.lm_ids if int(v) not in fixed] + [(int(v), 3) for v in g.pose_ids if int(v) not in fixed]
    ids.sort()
    bidx = {v: k for k, (v, _) in enumerate(ids)}
    dims = np.array([d for _, d in ids], dtype=np.int32)
    pairs = set()
    for a, b in zip(g.eo_from, g.eo_to):
        if int(a) in bidx and int(b) in bidx:
            x, y = bidx[int(a)], bidx[int(b)]
            pairs.add((min(x, y), max(x, y)))
    for a, b in zip(g.el_pose, g.el_lm):
        if int(a) in bidx and int(b) in bidx:
            x, y = bidx[int(a)], bidx[int(b)]
            pairs.add((min(x, y), max(x, y)))
    pairs = sorted(pairs)
    pa = np.array([p[0] for p in pairs], dtype=np.int32)
    pb = np.array([p[1] for p in pairs], dtype=np.int32)
    return ids, dims, pa, pb


def hvals_from_dense(H, dims, pa, pb, sym):
    """Block value array in the layout SymbolicAnalysis assumes, from a dense symmetric H (g2o order)."""
    off = np.concatenate([[0], np.cumsum(dims)])
    hv = np.zeros(sym.nH)
    for b in range(len(dims)):
        blk = H[off[b]:off[b + 1], off[b]:off[b + 1]]
        hv[sym.hoff_diag[b]:sym.hoff_diag[b] + blk.size] = blk.ravel()
    for k, (a, b) in enumerate(zip(pa, pb)):
        blk = H[off[a]:off[a + 1], off[b]:off[b + 1]]
        hv[sym.hoff_off[k]:sym.hoff_off[k] + blk.size] = blk.ravel()
    return hv


def factor_solve(sym, hv, b_solver):
    """Runs the front schedule.  b_solver: rhs in solver (permuted) scalar order.  Returns x (solver order)."""
    nf = sym.nf
    Ls, Us, uvec = [None] * nf, [None] * nf, [None] * nf
    x = np.zeros(sym.n)
    kids = [sym.children[sym.child_ptr[f]:sym.child_ptr[f + 1]] for f in range(nf)]
    for f in range(nf):  # fronts are numbered level-major: children always have smaller ids
        s, u = int(sym.npiv[f]), int(sym.nupd[f])
        fs = s + u
        F = np.zeros((fs, fs))
        for q in range(sym.asm_ptr[f], sym.asm_ptr[f + 1]):
            hoff, r, c, meta = (int(v) for v in sym.asm[q])
            dr, dc, trans, diag = meta & 0xff, (meta >> 8) & 0xff, (meta >> 16) & 1, (meta >> 17) & 1
            blk = hv[hoff:hoff + dr * dc].reshape(dr, dc)
            if diag:
                F[r:r + dr, c:c + dc] = np.tril(blk)
            elif not trans:
                F[r:r + dr, c:c + dc] = blk
            else:
                F[r:r + dc, c:c + dr] = blk.T
        for ch in kids[f]:
            assert ch < f
            rel = sym.rel[sym.rows_ptr[ch]:sym.rows_ptr[ch + 1]]
            assert np.all(rel >= 0) and np.all(rel < fs) and np.all(np.diff(rel) > 0)
            F[np.ix_(rel, rel)] += np.tril(Us[ch])
        for k in range(s):
            d = F[k, k]
            col = F[k + 1:, k].copy()
            F[k + 1:, k + 1:] -= np.tril(np.outer(col, col / d))
        L = np.tril(F[:, :s]).copy()
        for k in range(s):
            L[k + 1:, k] /= L[k, k]
        Ls[f] = L
        Us[f] = np.tril(F[s:, s:])
    # forward
    for f in range(nf):
        s, u = int(sym.npiv[f]), int(sym.nupd[f])
        p0 = int(sym.piv0[f])
        w = np.zeros(s + u)
        w[:s] = b_solver[p0:p0 + s]
        for ch in kids[f]:
            rel = sym.rel[sym.rows_ptr[ch]:sym.rows_ptr[ch + 1]]
            w[rel] += uvec[ch]
        L = Ls[f]
        for k in range(s):
            w[k + 1:] -= L[k + 1:, k] * w[k]
        x[p0:p0 + s] = w[:s] / np.diag(L)[:s]
        uvec[f] = w[s:]
    # backward
    for f in range(nf - 1, -1, -1):
        s, u = int(sym.npiv[f]), int(sym.nupd[f])
        p0 = int(sym.piv0[f])
        rows = sym.upd_rows[sym.rows_ptr[f]:sym.rows_ptr[f + 1]]
        xs = np.concatenate([x[p0:p0 + s], x[rows]])
        L = Ls[f]
        for k in range(s):
            xs[k] -= L[s:, k] @ xs[s:]
        for i in range(s - 1, 0, -1):
            xs[:i] -= L[i, :i] * xs[i]
        x[p0:p0 + s] = xs[:s]
    return x


def solver_perm(sym, dims):
    """perm[solver scalar index] = g2o scalar index."""
    off = np.concatenate([[0], np.cumsum(dims)])
    perm = np.zeros(sym.n, dtype=np.int64)
    for b in range(len(dims)):
        for k in range(dims[b]):
            perm[sym.boff[b] + k] = off[b] + k
    return perm

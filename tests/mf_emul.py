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


# ------------------------------------------------------------------------------------------------
# Tiled batched factorisation (csrc/tileplan.h, solver.cu: factor_tile_kernel / backward_tile_kernel)
# ------------------------------------------------------------------------------------------------
def _tile_off(i, j):
    I, J = i >> 3, j >> 3
    return ((I * (I + 1)) // 2 + J) * 64 + (i & 7) * 8 + ((j & 7) ^ ((i & 2) << 1))   # tileplan.h: tile_off


def tile_factor_solve(sym, hv, b_solver):
    """Emulates, front by front, exactly what one warp of the tiled kernels does with the host plan: zero the tile
    storage, identity on the padding pivots, add the plan's items (first from V = [H values | rhs], then from the
    children's stored fronts), eliminate the KT pivot tile columns of the augmented front (rhs = last local row),
    store the whole tile triangle, read z from the rhs row; then the backward sweep from the stored L tiles."""
    assert sym.tile_ok
    V = np.concatenate([hv, np.zeros(max(0, sym.tile_rhs_base - len(hv))), b_solver])
    Fv = np.zeros(sym.tile_nF)
    x = np.zeros(sym.n)
    nf = sym.nf

    def dense_from_tiles(st, nloc):
        A = np.zeros((nloc, nloc))
        for i in range(nloc):
            for j in range(i + 1):
                A[i, j] = st[_tile_off(i, j)]
        return A

    for f in range(nf):
        s, u = int(sym.npiv[f]), int(sym.nupd[f])
        sp = (s + 7) & ~7
        nloc = sp + u + 1
        T, KT = int(sym.tile_T[f]), int(sym.tile_KT[f])
        assert T == (nloc + 7) // 8 and KT == sp // 8
        ntile = T * (T + 1) // 2
        st = np.zeros(ntile * 64)
        for p in range(s, sp):
            st[_tile_off(p, p)] = 1.0
        a0, a1, nv = int(sym.tile_item_ptr[f]), int(sym.tile_item_ptr[f + 1]), int(sym.tile_item_nv[f])
        assert nv % 32 == 0 and (a1 - a0) % 32 == 0
        for g0 in range(a0, a1, 32):   # one warp step: 32 items, destinations must be distinct
            grp = sym.tile_items[g0:g0 + 32]
            live = grp[grp[:, 0] >= 0]
            assert len(set(live[:, 1].tolist())) == len(live), "two items of one warp step share a destination"
            src = V if g0 < a0 + nv else Fv
            for sidx, d in live:
                assert 0 <= d < ntile * 64
                st[d] += src[sidx]
        A = dense_from_tiles(st, nloc)           # lower triangle of the augmented local front
        for k in range(sp):                       # pivot columns (padding pivots: d = 1, column 0)
            d = A[k, k]
            col = A[k + 1:, k].copy()
            A[k + 1:, k + 1:] -= np.tril(np.outer(col, col / d))
            A[k + 1:, k] = col / d                # scaled: unit lower L, D stays on the diagonal
        for i in range(nloc):
            for j in range(i + 1):
                st[_tile_off(i, j)] = A[i, j]
        Fv[sym.tile_fptr[f]:sym.tile_fptr[f] + ntile * 64] = st
        p0 = int(sym.piv0[f])
        x[p0:p0 + s] = A[sp + u, :s]              # z = D^-1 L^-1 b sits in the rhs row
    for f in range(nf - 1, -1, -1):
        s, u = int(sym.npiv[f]), int(sym.nupd[f])
        sp = (s + 7) & ~7
        nloc = sp + u + 1
        st = Fv[sym.tile_fptr[f]:]
        p0 = int(sym.piv0[f])
        rows = sym.upd_rows[sym.rows_ptr[f]:sym.rows_ptr[f + 1]]
        xs = np.zeros(nloc)
        xs[:s] = x[p0:p0 + s]
        xs[sp:sp + u] = x[rows]
        Lm = dense_from_tiles(st, nloc)
        for K in range(sp // 8 - 1, -1, -1):
            k0 = 8 * K
            w = xs[k0:k0 + 8] - Lm[k0 + 8:, k0:k0 + 8].T @ xs[k0 + 8:]   # xs[rhs row] = 0: the z row drops out
            for i in range(7, -1, -1):
                for j in range(i):
                    w[j] -= Lm[k0 + i, k0 + j] * w[i]
            xs[k0:k0 + 8] = w
        x[p0:p0 + s] = xs[:s]
    return x

// tileplan.cpp -- see tileplan.h.  Pure host code (no CUDA): exported through the symbolic handle so the
// CPU test suite can run the plan through a numpy emulation of the kernel (tests/mf_emul.py).
#include "tileplan.h"

#include <climits>

namespace {
inline int tile_off_checked(int i, int j) { return i >= j ? tile_off(i, j) : -1; }
inline void pad32(std::vector<TileItem>& v, size_t group_start) {
  while ((v.size() - group_start) % 32) v.push_back({-1, 0});
}
}  // namespace

void tile_plan_build(const Symbolic& S, const std::vector<int>& solver2v, TilePlan& P) {
  P = TilePlan();
  P.nf = S.nf;
  P.T.resize(S.nf);
  P.KT.resize(S.nf);
  P.fptr.assign(S.nf + 1, 0);
  P.item_ptr.assign(S.nf + 1, 0);
  P.item_nv.assign(S.nf, 0);
  for (int f = 0; f < S.nf; f++) {
    const int s = S.npiv[f], u = S.nupd[f], sp = tile_sp(s), nloc = sp + u + 1;
    if (nloc > TILE_MAX_ROWS) return;  // not a tiled topology: the caller keeps the general path
    P.T[f] = (nloc + 7) / 8;
    P.KT[f] = sp / 8;
    P.max_T = P.T[f] > P.max_T ? P.T[f] : P.max_T;
    P.fptr[f + 1] = P.fptr[f] + (long)(P.T[f] * (P.T[f] + 1) / 2) * 64;
  }
  if (P.fptr[S.nf] > (long)INT_MAX) return;  // item sources are ints
  for (int f = 0; f < S.nf; f++) {
    const int s = S.npiv[f], u = S.nupd[f], sp = tile_sp(s), rhs = sp + u;
    auto loc = [&](int p) { return p < s ? p : sp + (p - s); };  // front index -> local row
    const size_t g0 = P.items.size();
    P.item_ptr[f] = (int)g0;
    // ---- H blocks (every block whose earlier-eliminated vertex is a pivot of this front) ----
    for (int q = S.asm_ptr[f]; q < S.asm_ptr[f + 1]; q++) {
      const AsmEntry& en = S.asm_entries[q];
      const int dr = en.meta & 0xff, dc = (en.meta >> 8) & 0xff;
      const bool trans = (en.meta >> 16) & 1, diag = (en.meta >> 17) & 1;
      // same three cases as the general kernels (solver.cu): block stored row-major dr x dc at hoff
      if (diag) {
        for (int i = 0; i < dr; i++)
          for (int j = 0; j <= i; j++) P.items.push_back({en.hoff + i * dc + j, tile_off_checked(loc(en.r + i), loc(en.c + j))});
      } else if (!trans) {
        for (int i = 0; i < dr; i++)
          for (int j = 0; j < dc; j++) P.items.push_back({en.hoff + i * dc + j, tile_off_checked(loc(en.r + i), loc(en.c + j))});
      } else {
        for (int i = 0; i < dc; i++)
          for (int j = 0; j < dr; j++) P.items.push_back({en.hoff + j * dc + i, tile_off_checked(loc(en.r + i), loc(en.c + j))});
      }
    }
    for (size_t q = g0; q < P.items.size(); q++)
      if (P.items[q].dst < 0) return;  // a block above the diagonal: not the orientation the fronts are assembled in
    // ---- right-hand side of the pivots into the rhs row ----
    for (int i = 0; i < s; i++) P.items.push_back({solver2v[S.piv0[f] + i], tile_off(rhs, i)});
    pad32(P.items, g0);
    P.item_nv[f] = (int)(P.items.size() - g0);
    // ---- children: Schur complement (lower triangle) and update vector, one padded group per child ----
    for (int ci = S.child_ptr[f]; ci < S.child_ptr[f + 1]; ci++) {
      const int ch = S.children[ci];
      const int uc = S.nupd[ch], spc = tile_sp(S.npiv[ch]);
      const int* rel = S.rel.data() + S.rows_ptr[ch];
      const long base = P.fptr[ch];
      const size_t c0 = P.items.size();
      for (int a = 0; a < uc; a++)
        for (int b = 0; b <= a; b++)
          P.items.push_back({(int)(base + tile_off(spc + a, spc + b)), tile_off(loc(rel[a]), loc(rel[b]))});
      for (int b = 0; b < uc; b++)  // the child's rhs row under its update columns = its update vector
        P.items.push_back({(int)(base + tile_off(spc + uc, spc + b)), tile_off(rhs, loc(rel[b]))});
      pad32(P.items, c0);
    }
  }
  P.item_ptr[S.nf] = (int)P.items.size();
  P.ok = true;
}

// tileplan.h -- host plan of the TILED batched factorisation (solver.cu: factor_tile_kernel).
//
// For a batch of replicas of one topology (config 3: thousands of Monte-Carlo replicas of the trackdrive
// graph, every front well under 96 rows) each (front, replica) is one warp and the dense front is kept as 8 x 8
// fp64 tiles, the shape of one mma.sync.m8n8k4.f64 accumulator.  This plan fixes, once per topology:
//   * the LOCAL layout of a front: pivots 0..s-1, identity padding up to sp = 8*ceil(s/8) (so that the
//     update rows start on a tile boundary), the u update rows, then ONE extra row carrying the right-hand
//     side (the forward solve rides along as a row of the augmented matrix); nloc = sp + u + 1 <= 96,
//     T = ceil(nloc / 8) tile rows, KT = sp / 8 pivot tile columns;
//   * the storage of a front, in shared memory while it is factorised and in HBM afterwards: the lower
//     triangle of tiles, tile (I, J), I >= J, at ((I (I+1)) / 2 + J) * 64 doubles, row-major inside a tile (with the
//     two column halves of rows 2, 3, 6, 7 swapped, see tile_off).
//     After the factorisation the tiles of columns J < KT hold L (unit lower, D on the diagonal; the rhs
//     row holds z = D^-1 L^-1 b), the tiles J >= KT hold the Schur complement and the update vector (rhs row);
//   * the assembly list of every front: {source offset, destination offset in the front} pairs, first the
//     entries of H and b (source = offset into the assembled system V), then, child by child, the entries
//     of the children's Schur complements and update vectors (source = offset into the replica's front
//     storage).  Every group is padded to a multiple of 32 items (source -1), so the 32 items a warp adds
//     in one step never share a destination.
// Replaces the per-entry index arithmetic (AsmEntry blocks, rel[] maps, packed-triangle offsets) the
// warp-per-front kernel redid for every replica: 12,000 warp instructions per front (ncu, round 1).
#pragma once
#include <vector>

#include "symbolic.h"

struct TileItem { int src, dst; };
constexpr int TILE_MAX_ROWS = 96;  // 12 tile rows: 78 tiles = 39 KB of shared memory per warp

struct TilePlan {
  bool ok = false;               // every front fits (nloc <= TILE_MAX_ROWS) and every offset fits an int
  int nf = 0;
  std::vector<int> T, KT;        // [nf]
  std::vector<long> fptr;        // [nf+1] offset of every front's tile storage (doubles) per replica
  std::vector<int> item_ptr;     // [nf+1] -> items
  std::vector<int> item_nv;      // [nf] leading items whose source is V (multiple of 32)
  std::vector<TileItem> items;
  int max_T = 0;
  long nF() const { return fptr.empty() ? 0 : fptr.back(); }
};

inline int tile_sp(int s) { return (s + 7) & ~7; }
// Local (i, j), i >= j -> offset inside the front's tile storage.  Inside a tile the rows with bit 1 set store
// their two 4-column halves swapped ((j & 7) ^ 4): the MMA A/B-operand reads (lane (g, t) takes column 4h + t of
// row g) then touch every shared-memory bank exactly twice (2 wavefronts, the minimum for 256 bytes) instead of
// four times, while the accumulator-layout reads (16 bytes per lane, two rows per 128 bytes) stay conflict-free.
inline int tile_off(int i, int j) {
  const int I = i >> 3, J = j >> 3;
  return ((I * (I + 1)) / 2 + J) * 64 + (i & 7) * 8 + ((j & 7) ^ ((i & 2) << 1));
}

// solver2v[k]: offset in V of the rhs entry of solver scalar k.
void tile_plan_build(const Symbolic& S, const std::vector<int>& solver2v, TilePlan& P);

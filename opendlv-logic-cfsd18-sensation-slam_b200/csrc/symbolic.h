// symbolic.h -- host symbolic phase of the block-sparse multifrontal LDL^T.
//
// Plays the role of Eigen's SimplicialLDLT::analyzePattern (thirdparty/Eigen/src/SparseCholesky/
// SimplicialCholesky.h:230-238: AMD ordering + elimination tree + column counts) that the reference
// reaches through g2o's LinearSolverEigen, re-designed for the GPU: a nested-dissection ordering of
// the vertex graph (short, wide assembly tree = few dependent kernel levels) whose tree nodes are
// the dense fronts the numeric kernels factorise.
#pragma once
#include <cstdint>
#include <functional>
#include <vector>

struct AsmEntry {   // one H block landing in a front's pivot columns
  int hoff;         // offset of the block in the H value array (row-major dr x dc as assembled)
  int r, c;         // top-left scalar position inside the front (r >= c)
  int meta;         // dr | dc << 8 | transposed << 16 | diagonal << 17   (dr, dc = dims as stored)
};

struct Symbolic {
  // ---- input echo ----
  int nb = 0;                       // free vertices (blocks), in g2o Hessian order (ascending id)
  int n = 0;                        // scalar dimension
  std::vector<int> dim;             // [nb] 2 or 3
  // ---- ordering ----
  std::vector<int> pos;             // [nb] elimination position of block b
  std::vector<int> boff;            // [nb] scalar offset of block b in solver (permuted) order
  // ---- fronts, children before parents; fronts of one level are contiguous ----
  int nf = 0, nlevels = 0;
  std::vector<int> level_ptr;       // [nlevels+1] -> front ids (fronts are numbered level-major)
  std::vector<int> piv0;            // [nf] first pivot scalar (solver order)
  std::vector<int> npiv, nupd;      // [nf] scalar counts
  std::vector<int> parent;          // [nf] or -1
  std::vector<int> rows_ptr;        // [nf+1] -> upd_rows / rel
  std::vector<int> upd_rows;        // global scalar index (solver order) of every update row
  std::vector<int> rel;             // position of every update row inside the parent front
  std::vector<int> child_ptr;       // [nf+1] -> children
  std::vector<int> children;
  std::vector<long> lptr;           // [nf+1] offsets of the L panels (fsize x npiv, column-major)
  std::vector<long> uptr;           // [nf+1] offsets of the update matrices (nupd x nupd)
  std::vector<int> asm_ptr;         // [nf+1] -> asm_entries
  std::vector<AsmEntry> asm_entries;
  // ---- statistics ----
  long nnzL = 0;                    // scalar entries of L below the diagonal (incl. explicit zeros)
  double flops = 0;                 // factor flops, sum over fronts
  int max_front = 0;
  double seconds = 0, t_nd = 0, t_md = 0;
};

// offdiag: nnb pairs (a, b), a < b, distinct, block indices in g2o order; hoff_diag[b] / hoff_off[k]
// give where the assembly kernels store each block (dims: diag d x d; off-diag dim[a] x dim[b]).
// leaf_size: nested dissection stops at subgraphs of at most this many vertices.
// amalgamate: 1 = latency-driven merging of critical-path children into their parents, 0 = none.
// Off by default: measured on B200 it shortens the assembly tree (10-lap graph 10 -> 8 levels, 1-lap
// graph 10 -> 5) but the GN iteration gets slower (507 -> 547 us, 315 -> 332 us): a panel of 8 pivots
// costs more than the fixed cost of a front, so fewer, larger fronts lose.  SLAM_B200_AMALG enables it.
// structure_ready (optional): called once, from the analysing thread, with true when everything but asm_ptr /
// asm_entries / seconds is final -- the caller may build what depends on the tree and the fronts (launch lists,
// gather lists) on another thread while the assembly entries are still being computed -- or with false on the way
// out of an exception thrown before that point.
void symbolic_analyze(int nb, const int* dim, int nnb, const int* off_a, const int* off_b,
                      const int* hoff_diag, const int* hoff_off, int leaf_size, Symbolic& S, int amalgamate = 0,
                      const std::function<void(bool)>* structure_ready = nullptr);

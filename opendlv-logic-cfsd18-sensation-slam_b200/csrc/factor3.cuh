// factor3.cuh -- included by solver.cu (inside its anonymous namespace, after the tile helpers).
//
// factor3_kernel: one CTA of 16 warps per LARGE front of a single graph (more than 64 rows, at most 160 local
// rows), the trailing matrix in REGISTERS.  factor2_kernel keeps the front in shared memory and its 8-wide trailing
// update reads and writes the whole trailing matrix once per panel: at the shared-memory bandwidth limit the
// 138-row root front of the 10-lap graph cannot go below ~55 us (measured 76 us).  Here the front is the tile
// triangle of the batched kernels (8 x 8 fp64 tiles; local layout: pivots | identity padding | update rows |
// right-hand side) and every tile lives in the registers of ONE warp in the accumulator layout of
// mma.sync.m8n8k4.f64; the 16 warps form a 4 x 4 grid, warp (a, b) owns the tiles (I, J) with I = a, J = b (mod 4),
// so per panel a warp fetches at most five row operands and five column operands from shared memory for up to
// fifteen tile updates (two DMMAs each).  Shared memory holds the assembly buffer, the finished L tiles (the
// operands of the updates) and the 8 x 8 inverse of the current diagonal tile.  Per pivot tile column: the owner of
// the diagonal tile factorises it by shuffles and publishes inverse + reciprocals | barrier | the four owners of
// the column's tiles turn them into L | barrier | every warp updates its tiles.  Inputs and outputs are exactly
// factor2_kernel's (assembly entries, children's Schur complements through rel[], gathered update vectors; L panels
// column-major, Schur complement, update vector, z), so the backward kernels and the parents do not notice which
// kernel produced a front.
//
// MEASURED (C2, B200, profiles/r02_factor3.md): correct (parity 1.6e-13, the whole GPU suite passes with it) but NOT
// faster than factor2_kernel -- 0.431 against 0.435 ms of factorisation per iteration, launch by launch within 5 %:
// the per-panel chain (one warp factorises the diagonal tile, four warps form the column's L tiles, two block
// barriers) costs what factor2_kernel's shared-memory traffic costs, and at 512 threads the kernel sits exactly at
// its 128-register limit.  It is therefore OFF by default (SLAM_B200_FACTOR_VARIANT=3 selects it); what it would
// need next is written down in DESIGN.md section 8.  Arithmetic: SimplicialCholesky_impl.h:122-191 (LDL^T without pivoting, failure iff a
// pivot is exactly zero), blocked by tiles.
#pragma once

constexpr int F3_THREADS = 512;
constexpr int F3_NI = 5;                       // tile rows (and columns) per warp
constexpr int F3_MAX_T = 4 * F3_NI;            // 20 tile rows = 160 local rows
constexpr int F3_SLOTS = (F3_NI * (F3_NI + 1)) / 2;

__host__ __device__ inline int f3_local_rows(int s, int u) { return ((s + 7) & ~7) + u + 1; }
inline size_t f3_smem_bytes(int nloc) {
  const int T = (nloc + 7) >> 3;
  return sizeof(double) * ((size_t)((T * (T + 1)) / 2) * 64 + 64 + 16) + sizeof(int) * (size_t)((nloc + 1) & ~1);
}

__global__ void __launch_bounds__(F3_THREADS)
factor3_kernel(SymArgs S, int list_off, const double* __restrict__ V_all, long nV, double* Lv_all, long nL,
               double* Uv_all, long nU, int* status, double* uvec_all, long nUvec, double* x_all, int n) {
  extern __shared__ double smem[];
  const int f = S.launch_list[list_off + blockIdx.x];
  const int r = blockIdx.y;
  const int s = S.npiv[f], u = S.nupd[f], fs = s + u;
  const int sp = (s + 7) & ~7, KT = sp >> 3, rr = sp + u, nloc = rr + 1, T = (nloc + 7) >> 3;
  const int NT = (T * (T + 1)) >> 1;
  const double* V = V_all + (size_t)r * nV;
  double* Uv = Uv_all + (size_t)r * nU;
  double* F = smem;
  double* Wb = F + (size_t)NT * 64;     // inverse of the current unit triangle, row-major 8 x 8
  double* dinvb = Wb + 64;              // 8 reciprocals of the current pivots (+ 8 spare)
  int* srel = reinterpret_cast<int*>(dinvb + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto loc = [&](int p) { return p < s ? p : sp + (p - s); };
  const bool dbgc = S.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;  // SLAM_B200_PHASE_CLOCKS
  if (dbgc) { S.dbg[0] = clock64(); for (int q = 24; q < 32; q++) S.dbg[q] = 0; }
  long long dt_ = 0, ck1 = 0, ck2 = 0, ck3 = 0, ck4 = 0, ck5 = 0;  // panel-step clocks, in registers until the end
  // ---- assembly in shared memory ----
  {
    double2* F2 = reinterpret_cast<double2*>(F);
    for (int q = tid; q < (NT << 5); q += F3_THREADS) F2[q] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  if (dbgc) S.dbg[1] = clock64();
  if (tid < sp - s) F[tile_at(s + tid, s + tid)] = 1.0;  // padding pivots: identity
  for (int q = S.asm_ptr[f] + tid; q < S.asm_ptr[f + 1]; q += F3_THREADS) {
    const AsmEntry en = S.asm_entries[q];
    const double* hv = V + en.hoff;
    const int dr = en.meta & 0xff, dc = (en.meta >> 8) & 0xff;
    const bool trans = (en.meta >> 16) & 1, diag = (en.meta >> 17) & 1;
    if (diag) {
      for (int i = 0; i < dr; i++)
        for (int j = 0; j <= i; j++) F[tile_at(loc(en.r + i), loc(en.c + j))] = hv[i * dc + j];
    } else if (!trans) {
      for (int i = 0; i < dr; i++)
        for (int j = 0; j < dc; j++) F[tile_at(loc(en.r + i), loc(en.c + j))] = hv[i * dc + j];
    } else {
      for (int i = 0; i < dc; i++)
        for (int j = 0; j < dr; j++) F[tile_at(loc(en.r + i), loc(en.c + j))] = hv[j * dc + i];
    }
  }
  {  // right-hand side row: rhs of the pivots + the children's update vectors, fixed (child) order
    const int p0 = S.piv0[f];
    const double* uvecr = uvec_all + (size_t)r * nUvec;
    const int* gp = S.gather_ptr + S.frow_ptr[f];
    for (int i = tid; i < fs; i += F3_THREADS) {
      double acc = i < s ? V[S.solver2v[p0 + i]] : 0.0;
      for (int q = gp[i]; q < gp[i + 1]; q++) acc += uvecr[S.gather_src[q]];
      F[tile_at(rr, loc(i))] = acc;
    }
  }
  __syncthreads();
  if (dbgc) S.dbg[2] = clock64();
  for (int ci = S.child_ptr[f]; ci < S.child_ptr[f + 1]; ci++) {  // extend-add, child after child (fixed order)
    const int ch = S.children[ci];
    const int uc = S.nupd[ch];
    const double* __restrict__ Uc = Uv + S.uptr[ch];
    const int* rel = S.rel + S.rows_ptr[ch];
    for (int i = tid; i < uc; i += F3_THREADS) srel[i] = loc(rel[i]);
    __syncthreads();
    // two columns per warp step, four row chunks each, all loads before any update (the phase is a chain of
    // global-memory round trips otherwise)
    for (int j = warp; j < uc; j += 2 * (F3_THREADS / 32)) {
      const int jb = j + F3_THREADS / 32;
      const bool hb = jb < uc;
      const double* colA = Uc + (size_t)j * uc;
      const double* colB = Uc + (size_t)(hb ? jb : j) * uc;
      const int cA = srel[j], cB = srel[hb ? jb : j];
      for (int i0 = 0; i0 < uc - j; i0 += 128) {
        double va[4], vb[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int ia = j + i0 + 32 * c + lane, ib = jb + i0 + 32 * c + lane;
          va[c] = ia < uc ? __ldg(colA + ia) : 0.0;
          vb[c] = (hb && ib < uc) ? __ldg(colB + ib) : 0.0;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int ia = j + i0 + 32 * c + lane, ib = jb + i0 + 32 * c + lane;
          if (ia < uc) {
            const int ri = srel[ia];
            F[ri >= cA ? tile_at(ri, cA) : tile_at(cA, ri)] += va[c];
          }
          if (hb && ib < uc) {
            const int ri = srel[ib];
            F[ri >= cB ? tile_at(ri, cB) : tile_at(cB, ri)] += vb[c];
          }
        }
      }
    }
    __syncthreads();
  }
  if (dbgc) S.dbg[3] = clock64();
  // ---- the tile triangle into registers: warp (a, b) takes I = a + 4 (i + [b > a]), J = b + 4 j, j <= i ----
  const int wa = warp >> 2, wb = warp & 3, sh = wb > wa ? 1 : 0;
  const int g = lane >> 2, t = lane & 3;
  const int sw = (g & 2) << 1;
  const int cl = g * 8 + ((2 * t) ^ sw), al0 = g * 8 + (t ^ sw), al1 = g * 8 + ((4 + t) ^ sw);
  double2 c[F3_SLOTS];
#pragma unroll
  for (int i = 0; i < F3_NI; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      const int I = wa + 4 * (i + sh), J = wb + 4 * j;
      c[(i * (i + 1)) / 2 + j] = (I < T && J <= I) ? *reinterpret_cast<const double2*>(F + tile_base(I, J) + cl)
                                                    : make_double2(0.0, 0.0);
    }
  __syncthreads();
  bool bad = false;
  for (int K = 0; K < KT; K++) {
    double* Dk = F + tile_base(K, K);
    if (dbgc) dt_ = clock64();
    // (1) the owner of the diagonal tile factorises it and publishes the inverse of its unit triangle + 1/d
    if (wa == (K & 3) && wb == (K & 3)) {
      const int ik = K >> 2;  // sh == 0 on the diagonal of the warp grid
#pragma unroll
      for (int i = 0; i < F3_NI; i++) {
        if (i == ik) {
          double di0 = 1.0, di1 = 1.0;
          double2& d = c[(i * (i + 1)) / 2 + i];
          bad |= tile_ldlt(d.x, d.y, di0, di1, g, t);
          *reinterpret_cast<double2*>(Dk + cl) = d;
          if (g == 0) { dinvb[2 * t] = di0; dinvb[2 * t + 1] = di1; }
        }
      }
      __syncwarp();
      double x[8];
#pragma unroll
      for (int j = 7; j >= 0; j--) {
        double acc = (j == g) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 7; k > j; k--) acc -= x[k] * Dk[tile_in(k, j)];
        x[j] = acc;
      }
      if (t == 0) {
#pragma unroll
        for (int k = 0; k < 8; k++) Wb[g * 8 + k] = x[k];
      }
    }
    if (dbgc) { const long long t_ = clock64(); ck1 += t_ - dt_; dt_ = t_; }
    __syncthreads();
    if (dbgc) { const long long t_ = clock64(); ck2 += t_ - dt_; dt_ = t_; }
    // (2) the owners of the column's tiles: X = A W (= L D), L = X D^-1, final; published as update operands
    if (wb == (K & 3)) {
      const double w0 = Wb[g * 8 + t], w1 = Wb[g * 8 + 4 + t];
      const double di0 = dinvb[2 * t], di1 = dinvb[2 * t + 1];
      const int jk = K >> 2;
#pragma unroll
      for (int i = 0; i < F3_NI; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
          const int I = wa + 4 * (i + sh);
          if (j == jk && I > K && I < T) {  // warp-uniform
            double* P = F + tile_base(I, K);
            double2& q = c[(i * (i + 1)) / 2 + j];
            *reinterpret_cast<double2*>(P + cl) = q;  // accumulator layout -> A-operand layout through the tile's slot
            __syncwarp();
            const double pa0 = P[al0], pa1 = P[al1];
            double x0 = 0.0, x1 = 0.0;
            dmma_acc(x0, x1, pa0, w0);
            dmma_acc(x0, x1, pa1, w1);
            q = make_double2(x0 * di0, x1 * di1);
            __syncwarp();
            *reinterpret_cast<double2*>(P + cl) = q;
          }
        }
    }
    if (dbgc) { const long long t_ = clock64(); ck3 += t_ - dt_; dt_ = t_; }
    __syncthreads();
    if (dbgc) { const long long t_ = clock64(); ck4 += t_ - dt_; dt_ = t_; }
    // (3) every warp updates its tiles right of the column: C_IJ -= (L_IK D) L_JK^T
    {
      const double nd0 = -Dk[tile_in(t, t)], nd1 = -Dk[tile_in(4 + t, 4 + t)];
      double la0[F3_NI], la1[F3_NI], lb0[F3_NI], lb1[F3_NI];
      const int k64 = K << 6;
#pragma unroll
      for (int i = 0; i < F3_NI; i++) {
        const int I = wa + 4 * (i + sh), J = wb + 4 * i;
        const bool vi = (I > K) & (I < T), vj = (J > K) & (J < T);
        const int oi = (((I * (I + 1)) >> 1) << 6) + k64, oj = (((J * (J + 1)) >> 1) << 6) + k64;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        if (vi) { a0 = F[oi + al0]; a1 = F[oi + al1]; }
        if (vj) { b0 = F[oj + al0]; b1 = F[oj + al1]; }
        la0[i] = a0 * nd0;
        la1[i] = a1 * nd1;
        lb0[i] = b0;
        lb1[i] = b1;
      }
      // (both k-halves of a tile back to back: issuing all first halves before any second one -- the two MMAs of a
      // tile depend on each other -- was measured and lost: it pushes the kernel over its 128 registers, 56 bytes
      // of spills, panels 114k -> 167k cycles on the root front)
#pragma unroll
      for (int i = 0; i < F3_NI; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
          const int I = wa + 4 * (i + sh), J = wb + 4 * j;
          if (J > K && J <= I && I < T) {  // warp-uniform
            double2& q = c[(i * (i + 1)) / 2 + j];
            dmma_acc(q.x, q.y, la0[i], lb0[j]);
            dmma_acc(q.x, q.y, la1[i], lb1[j]);
          }
        }
    }
    if (dbgc) { const long long t_ = clock64(); ck5 += t_ - dt_; dt_ = t_; }
    // no barrier here: step (1) of the next column touches the owner's registers, its own diagonal slot and the
    // W / 1/d buffers, which nobody reads after the barrier in front of step (3)
  }
  if (dbgc) { S.dbg[4] = clock64(); S.dbg[5] = S.dbg[4]; }
  // ---- Schur complement, update vector and z back through shared memory, then out in factor2_kernel's formats ----
#pragma unroll
  for (int i = 0; i < F3_NI; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      const int I = wa + 4 * (i + sh), J = wb + 4 * j;
      if (J >= KT && J <= I && I < T) *reinterpret_cast<double2*>(F + tile_base(I, J) + cl) = c[(i * (i + 1)) / 2 + j];
    }
  if (bad && lane == 0) status[2 * r] = 1;
  __syncthreads();
  {
    const int p0 = S.piv0[f];
    double* xr = x_all + (size_t)r * n;
    double* uo = uvec_all + (size_t)r * nUvec + S.rows_ptr[f];
    for (int i = tid; i < fs; i += F3_THREADS) {
      const double v = F[tile_at(rr, loc(i))];
      if (i < s) xr[p0 + i] = v;  // the rhs row was scaled with the L rows: z = D^-1 L^-1 b
      else uo[i - s] = v;
    }
  }
  double* Lg = Lv_all + (size_t)r * nL + S.lptr[f];
  for (int j = warp; j < s; j += F3_THREADS / 32) {
    double* out = Lg + (size_t)j * fs;
    for (int i = lane; i < fs; i += 32) out[i] = i < j ? 0.0 : F[tile_at(loc(i), j)];
  }
  double* Ug = Uv + S.uptr[f];
  for (int j = warp; j < u; j += F3_THREADS / 32) {
    double* out = Ug + (size_t)j * u;
    for (int i = j + lane; i < u; i += 32) out[i] = F[tile_at(sp + i, sp + j)];
  }
  __syncthreads();
  if (dbgc) { S.dbg[24] = ck1; S.dbg[25] = ck2; S.dbg[26] = ck3; S.dbg[27] = ck4; S.dbg[28] = ck5; }
  if (dbgc) { S.dbg[6] = clock64(); S.dbg[7] = s; S.dbg[8] = fs; S.dbg[9] = S.child_ptr[f + 1] - S.child_ptr[f]; }
}

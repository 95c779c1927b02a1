// solver.cu -- numeric phase of the block-sparse multifrontal LDL^T and the triangular solves
// (sm_100a, fp64).
//
// Replaces Eigen::SimplicialLDLT::factorize / _solve_impl (thirdparty/Eigen/src/SparseCholesky/
// SimplicialCholesky_impl.h:101-195, SimplicialCholesky.h:156-180) that the reference reaches via
// g2o's LinearSolverEigen::solve from Slam::optimizeGraph (slam.cpp:481).  Same mathematics
// (H = L D L^T without pivoting, failure iff a pivot is exactly zero), different schedule: the
// assembly tree of symbolic.cpp is walked level by level; every front is one CTA that assembles
// its dense frontal matrix in shared memory (original H blocks + the children's Schur complements),
// eliminates its pivot columns and hands its own Schur complement to the parent.  Landmark and pose
// blocks are ordered inside this factorisation; each update matrix IS the Schur complement of the
// eliminated variables, it is just never formed globally.
#include <array>
#include <cstring>

#include "graph_dev.h"

namespace {

constexpr int FACTOR_THREADS = 512;

constexpr int SOLVE_THREADS = 1024;      // levels with fronts of > 64 rows: few CTAs, stage L fast


struct SymArgs {
  const int *piv0, *npiv, *nupd, *rows_ptr, *upd_rows, *rel, *child_ptr, *children, *asm_ptr, *solver2v;
  const int *frow_ptr, *gather_ptr, *gather_src;
  long long* dbg;  // optional phase clocks of the last-launched CTA-per-front kernel (SLAM_B200_PHASE_CLOCKS)
  int dbg_front;   // front whose CTA records the phase clocks (-1: block 0 of every launch, the last one wins)
  long long* tl;   // optional front timeline (SLAM_B200_TIMELINE): per front %globaltimer at factor entry / after the
                   // wait / end, backward entry / after the wait / end
  const long *lptr, *uptr, *fbig;
  const AsmEntry* asm_entries;
  const int* launch_list;
};

// ---- factorisation -----------------------------------------------------------------------------
// F is the dense frontal matrix, column-major with leading dimension fs; only the lower triangle
// is meaningful.  SMEM: F in dynamic shared memory; otherwise in a per-front global scratch slab.
//
// Blocked right-looking LDL^T, panels of NB pivot columns:
//  (1) panel: every thread factorises the NB x NB diagonal triangle redundantly in registers (it is
//      tiny and this avoids a barrier + broadcast), then one thread per row below it eliminates
//      that row's NB panel entries in registers;
//  (2) trailing update F[i,j] -= sum_p F[i,p] F[j,p] / d_p: one warp per group of 4 columns, lanes
//      over rows, the 4 x NB scaled panel entries of the columns held in registers -> 32 FMAs per 8
//      shared-memory loads, conflict-free (lanes read consecutive rows).
// Two barriers per NB pivots.  Columns stay unscaled (F[i,k] = l_ik d_k) until the write-out.
constexpr int NB = 8;

// ---- programmatic dependent launch (single-graph path) ---------------------------------------------
// The levels of the assembly tree are dependent launches of 7-70 us each; a quarter of a level is spent before the
// first byte of the previous level's output is needed (launch latency, the index chain launch list -> front sizes ->
// assembly entries, zeroing the front, scattering the H blocks / staging the L panel).  The front kernels are
// therefore launched with cudaLaunchAttributeProgrammaticStreamSerialization and follow ONE protocol:
//   prologue (reads only the static structure and data of kernels at least TWO launches back) ->
//   griddepcontrol.wait (the previous launch has completed, its writes are visible) ->
//   griddepcontrol.launch_dependents (the next launch may become resident and run its prologue) -> body.
// Because a CTA releases its dependents only after its own wait, a prologue never overlaps anything older than the
// launch directly before it: what it reads (V from the assembly kernels, L from the factor kernels) is complete unless
// that launch is the producer itself -- the first factor launch (V) and the first backward launch (L) get early = 0
// and wait before they touch anything.  Data of the previous launch (children's Schur complements and update
// vectors, the parents' x) is read after the wait with ld.global.cg.  Without the launch attribute both
// instructions are no-ops.  SLAM_B200_NO_PDL=1 launches everything the plain way (A/B measurements).
constexpr int TLS = 12;  // stamps per front in the timeline: factor entry / wait passed / rhs + row maps staged /
                         // children added / panels done / end, backward entry / wait passed / end, 3 spare
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// D = A B + C on the fp64 tensor pipe: A 8 x 4 (lane l: row l / 4, column l % 4), B 4 x 8 (lane l: row l % 4, column
// l / 4), C / D 8 x 8 (lane l: row l / 4, columns 2 (l % 4) and 2 (l % 4) + 1)
__device__ __forceinline__ void dmma_upd(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// LDL^T of an 8 x 8 tile in the accumulator layout (lane (g, t) = (l / 4, l % 4): entries (g, 2t), (g, 2t + 1)), by
// shuffles and without redundancy -- the routine of the batched tile kernels (tile_ldlt below) with the two
// conventions of the CTA-per-front kernel: columns stay UNSCALED (l_qp d_p below the diagonal, d_p on it) and only an
// exactly zero pivot among the first nb fails.  di0 / di1 = reciprocals of the pivots of the lane's two columns.
__device__ __forceinline__ bool tile_ldlt_unscaled(double& a0, double& a1, double& di0, double& di1, int g, int t, int nb) {
  bool bad = false;
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int hp = p >> 1;
    const double v = (p & 1) ? a1 : a0;
    const double rgp = __shfl_sync(0xffffffffu, v, g * 4 + hp);      // T[g][p]
    const double dp = __shfl_sync(0xffffffffu, v, p * 4 + hp);       // T[p][p]
    const double c0 = __shfl_sync(0xffffffffu, v, 8 * t + hp);       // T[2t][p]
    const double c1 = __shfl_sync(0xffffffffu, v, 8 * t + 4 + hp);   // T[2t+1][p]
    bad |= (p < nb && dp == 0.0);
    double inv;
    {  // hardware seed + two Newton steps: within an ulp or two of 1 / dp (fast_rcp below)
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(dp));
      double e = fma(-dp, inv, 1.0);
      inv = fma(inv, e, inv);
      e = fma(-dp, inv, 1.0);
      inv = fma(inv, e, inv);
    }
    const double lgp = rgp * inv;
    if (2 * t > p) a0 -= lgp * c0;
    if (2 * t + 1 > p) a1 -= lgp * c1;
    if (t == hp) {
      if (p & 1) di1 = inv;
      else di0 = inv;
    }
  }
  return bad;
}

__device__ __forceinline__ void pdl_wait_then_release() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <bool SMEM>
__global__ void __launch_bounds__(FACTOR_THREADS)
factor_kernel(SymArgs S, int list_off, const double* __restrict__ V_all, long nV, double* Lv_all, long nL,
              double* Uv_all, long nU, double* Fbig_all, long nFbig, int* status, double* uvec_all, long nUvec,
              double* x_all, int n) {
  extern __shared__ double smem[];
  const int g = S.launch_list[list_off + blockIdx.x];
  const int r = blockIdx.y;
  const int s = S.npiv[g], u = S.nupd[g], fs = s + u;
  const double* V = V_all + (size_t)r * nV;
  double* Uv = Uv_all + (size_t)r * nU;
  double* F = SMEM ? smem : (Fbig_all + (size_t)r * nFbig + S.fbig[g]);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  if (S.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) S.dbg[0] = clock64();
  for (int t = tid; t < fs * fs; t += nt) F[t] = 0.0;
  __syncthreads();
  if (S.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) S.dbg[1] = clock64();
  // original entries: every H block whose earlier-eliminated vertex is a pivot of this front
  for (int q = S.asm_ptr[g] + tid; q < S.asm_ptr[g + 1]; q += nt) {
    const AsmEntry en = S.asm_entries[q];
    const double* hv = V + en.hoff;
    const int dr = en.meta & 0xff, dc = (en.meta >> 8) & 0xff;
    const bool trans = (en.meta >> 16) & 1, diag = (en.meta >> 17) & 1;
    if (diag) {
      for (int i = 0; i < dr; i++)
        for (int j = 0; j <= i; j++) F[(size_t)(en.c + j) * fs + en.r + i] = hv[i * dc + j];
    } else if (!trans) {
      for (int i = 0; i < dr; i++)
        for (int j = 0; j < dc; j++) F[(size_t)(en.c + j) * fs + en.r + i] = hv[i * dc + j];
    } else {
      for (int i = 0; i < dc; i++)
        for (int j = 0; j < dr; j++) F[(size_t)(en.c + j) * fs + en.r + i] = hv[j * dc + i];
    }
  }
  __syncthreads();
  // extend-add the children's Schur complements (one child at a time: positions of different
  // children overlap, positions inside one child do not); warp per column, lanes over rows
  if (S.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) S.dbg[2] = clock64();
  int* srel = reinterpret_cast<int*>(smem + (SMEM ? (size_t)fs * fs : 0));  // fs ints behind the front
  for (int ci = S.child_ptr[g]; ci < S.child_ptr[g + 1]; ci++) {
    const int ch = S.children[ci];
    const int uc = S.nupd[ch];
    const double* Uc = Uv + S.uptr[ch];
    const int* rel = S.rel + S.rows_ptr[ch];
    for (int i = tid; i < uc; i += nt) srel[i] = rel[i];
    __syncthreads();
    for (int j = warp; j < uc; j += nw) {
      const double* col = Uc + (size_t)j * uc;
      double* dst = F + (size_t)srel[j] * fs;
      for (int i = j + lane; i < uc; i += 128) {  // four independent loads in flight per lane
        const int i1 = i + 32, i2 = i + 64, i3 = i + 96;
        const double v0 = col[i];
        const double v1 = i1 < uc ? col[i1] : 0.0;
        const double v2 = i2 < uc ? col[i2] : 0.0;
        const double v3 = i3 < uc ? col[i3] : 0.0;
        dst[srel[i]] += v0;
        if (i1 < uc) dst[srel[i1]] += v1;
        if (i2 < uc) dst[srel[i2]] += v2;
        if (i3 < uc) dst[srel[i3]] += v3;
      }
    }
    __syncthreads();
  }
  // Sp: the current panel scaled by 1/d (Sp[p*fs + j] = F[j, k0+p] / d_p), written by the row threads
  // of step (1) and broadcast-read as the column factors of step (2)
  if (S.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) S.dbg[3] = clock64();
  double* Sp = reinterpret_cast<double*>(srel + ((fs + 1) & ~1));
  double* dinv = Sp + (size_t)NB * fs;  // 1/d of every pivot (s doubles), then w (fs) for the fused forward solve
  double* Tsm = dinv + 2 * (size_t)fs;   // current panel: NB x NB triangle + NB reciprocals, then NB forward values
  const bool dbgc = S.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;  // finer clocks, thread 0 (in warp 0)
  long long dt_ = 0;
  if (dbgc) for (int k = 24; k < 32; k++) S.dbg[k] = 0;  // like the stamps: the last launch (the root) wins
  for (int k0 = 0; k0 < s; k0 += NB) {
    const int nb = min(NB, s - k0);
    double* Pk = F + (size_t)k0 * fs;  // panel columns: Pk[p * fs + row]
    if (dbgc) dt_ = clock64();
    // ---- (1a) warp 0 factorises the NB x NB diagonal triangle in registers (a dependent chain of
    // 8 reciprocals: one warp, not sixteen, pays its instruction stream) and publishes the final
    // triangle + reciprocals of the pivots in shared memory
    if (warp == 0) {
      double T[NB][NB], iv[NB];
#pragma unroll
      for (int p = 0; p < NB; p++)
#pragma unroll
        for (int q = p; q < NB; q++) T[q][p] = (q < nb) ? Pk[p * fs + k0 + q] : (q == p ? 1.0 : 0.0);
      bool bad = false;
#pragma unroll
      for (int p = 0; p < NB; p++) {
        const double d = T[p][p];
        if (p < nb && d == 0.0) bad = true;  // SimplicialCholesky_impl.h:175-179: an exactly zero pivot fails, NaN / inf flow on
        iv[p] = __drcp_rn(d);
#pragma unroll
        for (int q = p + 1; q < NB; q++) {
          const double lqp = T[q][p] * iv[p];
#pragma unroll
          for (int q2 = q; q2 < NB; q2++) T[q2][q] -= T[q2][p] * lqp;
        }
      }
      if (lane == 0) {
        if (bad) status[2 * r] = 1;
#pragma unroll
        for (int p = 0; p < NB; p++) {
          Tsm[NB * NB + p] = iv[p];
          if (p < nb) dinv[k0 + p] = iv[p];
#pragma unroll
          for (int q = p; q < NB; q++) {
            Tsm[q * NB + p] = T[q][p];
            if (q < nb && p < nb) Pk[p * fs + k0 + q] = T[q][p];  // the triangle's own rows
          }
        }
      }
    }
    if (dbgc) { const long long t_ = clock64(); S.dbg[24] += t_ - dt_; dt_ = t_; }
    __syncthreads();
    if (dbgc) { const long long t_ = clock64(); S.dbg[25] += t_ - dt_; dt_ = t_; }
    // ---- (1b) one thread per row below the triangle eliminates that row's panel entries ----
    double invd[NB];
#pragma unroll
    for (int p = 0; p < NB; p++) invd[p] = Tsm[NB * NB + p];
    if (k0 + nb + tid < fs) {
      double T[NB][NB];
#pragma unroll
      for (int p = 0; p < NB; p++)
#pragma unroll
        for (int q = p + 1; q < NB; q++) T[q][p] = Tsm[q * NB + p];
      for (int i = k0 + nb + tid; i < fs; i += nt) {
        double rr[NB];
#pragma unroll
        for (int p = 0; p < NB; p++) rr[p] = (p < nb) ? Pk[p * fs + i] : 0.0;
#pragma unroll
        for (int p = 0; p < NB; p++) {
          const double rp = rr[p] * invd[p];
          Sp[p * fs + i] = rp;
#pragma unroll
          for (int q = p + 1; q < NB; q++) rr[q] -= rp * T[q][p];
        }
#pragma unroll
        for (int p = 1; p < NB; p++)
          if (p < nb) Pk[p * fs + i] = rr[p];
      }
    }
    if (dbgc) { const long long t_ = clock64(); S.dbg[26] += t_ - dt_; dt_ = t_; }
    __syncthreads();
    if (dbgc) { const long long t_ = clock64(); S.dbg[27] += t_ - dt_; dt_ = t_; }
    // ---- (2) trailing update ----
    const int c0 = k0 + nb;
    for (int jg = c0 + 4 * warp; jg < fs; jg += 4 * nw) {
      double B[4][NB];
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int p = 0; p < NB; p++) B[b][p] = (jg + b < fs) ? Sp[p * fs + jg + b] : 0.0;
      double* Cj = F + (size_t)jg * fs;
      for (int i = jg + lane; i < fs; i += 64) {  // two independent row chunks in flight
        const int i2 = i + 32;
        const bool v2 = i2 < fs;
        double A[NB], A2[NB];
#pragma unroll
        for (int p = 0; p < NB; p++) {
          A[p] = (p < nb) ? Pk[p * fs + i] : 0.0;
          A2[p] = (p < nb && v2) ? Pk[p * fs + i2] : 0.0;
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
          if (jg + b < fs) {
            const bool w1 = i >= jg + b, w2 = v2;  // i2 > i >= jg, so i2 >= jg + b always holds for b < 32
            double acc = w1 ? Cj[b * fs + i] : 0.0;
            double acc2 = w2 ? Cj[b * fs + i2] : 0.0;
#pragma unroll
            for (int p = 0; p < NB; p++) {
              acc -= A[p] * B[b][p];
              acc2 -= A2[p] * B[b][p];
            }
            if (w1) Cj[b * fs + i] = acc;
            if (w2) Cj[b * fs + i2] = acc2;
          }
        }
      }
    }
    if (dbgc) { const long long t_ = clock64(); S.dbg[28] += t_ - dt_; dt_ = t_; }
    __syncthreads();
    if (dbgc) { const long long t_ = clock64(); S.dbg[29] += t_ - dt_; dt_ = t_; }
  }
  if (S.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) S.dbg[4] = clock64();
  // ---- fused forward solve of this front (L y = b, z = D^-1 y): the factor is still in shared
  // memory, so the leaves-to-root sweep rides along with the factorisation instead of re-reading
  // every L panel in a second pass of kernels.  Columns are unscaled: l_ik = F[i,k] * dinv[k].
  {
    double* w = dinv + s;  // fs
    const int p0 = S.piv0[g];
    const double* uvecr = uvec_all + (size_t)r * nUvec;
    const int* gp = S.gather_ptr + S.frow_ptr[g];
    for (int i = tid; i < fs; i += nt) {
      double acc = i < s ? V[S.solver2v[p0 + i]] : 0.0;
      for (int q = gp[i]; q < gp[i + 1]; q++) acc += uvecr[S.gather_src[q]];
      w[i] = acc;
    }
    __syncthreads();
    double* xr = x_all + (size_t)r * n;
    for (int k0 = 0; k0 < s; k0 += NB) {
      const int nb = min(NB, s - k0);
      const double* Pk = F + (size_t)k0 * fs;
      double* ysm = Tsm + NB * NB + NB;  // NB forward values of this panel (z = y / d)
      if (dbgc) dt_ = clock64();
      if (warp == 0) {
        double yd[NB];
#pragma unroll
        for (int p = 0; p < NB; p++) yd[p] = (p < nb) ? w[k0 + p] : 0.0;
#pragma unroll
        for (int p = 0; p < NB; p++) {
          if (p < nb) {
            yd[p] *= dinv[k0 + p];
#pragma unroll
            for (int q = p + 1; q < NB; q++)
              if (q < nb) yd[q] -= Pk[p * fs + k0 + q] * yd[p];
          }
        }
        if (lane == 0) {
#pragma unroll
          for (int p = 0; p < NB; p++) {
            ysm[p] = yd[p];
            if (p < nb) xr[p0 + k0 + p] = yd[p];  // z = D^-1 y
          }
        }
      }
      if (dbgc) { const long long t_ = clock64(); S.dbg[30] += t_ - dt_; dt_ = t_; }
      __syncthreads();
      for (int i = k0 + nb + tid; i < fs; i += nt) {
        double acc = w[i];
#pragma unroll
        for (int p = 0; p < NB; p++)
          if (p < nb) acc -= Pk[p * fs + i] * ysm[p];
        w[i] = acc;
      }
      __syncthreads();
      if (dbgc) { const long long t_ = clock64(); S.dbg[31] += t_ - dt_; dt_ = t_; }
    }
    double* uo = uvec_all + (size_t)r * nUvec + S.rows_ptr[g];
    for (int i = s + tid; i < fs; i += nt) uo[i - s] = w[i];
  }
  if (S.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) S.dbg[5] = clock64();
  // L panel (fs x s, unit lower with D on the diagonal) and the Schur complement for the parent
  double* Lg = Lv_all + (size_t)r * nL + S.lptr[g];
  for (int j = warp; j < s; j += nw) {
    const double* col = F + (size_t)j * fs;
    const double d = col[j];
    const double inv = 1.0 / d;
    double* out = Lg + (size_t)j * fs;
    for (int i = lane; i < fs; i += 32) out[i] = i < j ? 0.0 : (i == j ? d : col[i] * inv);
  }
  double* Ug = Uv + S.uptr[g];
  for (int j = warp; j < u; j += nw) {
    const double* col = F + (size_t)(s + j) * fs + s;
    double* out = Ug + (size_t)j * u;
    for (int i = j + lane; i < u; i += 32) out[i] = col[i];
  }
  __syncthreads();
  if (S.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) S.dbg[6] = clock64();
  if (S.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) { S.dbg[7] = s; S.dbg[8] = fs; S.dbg[9] = S.child_ptr[g + 1] - S.child_ptr[g]; }
}

// LDL^T of one NB x NB diagonal triangle held (redundantly) in the registers of every lane of a warp;
// publishes the factorised triangle (unscaled: t_qp = l_qp d_p, d on the diagonal) and the reciprocals
// of the pivots to Tout[NB*NB + NB].  All lanes store the same values to the same addresses (one
// wavefront per store, no lane predicate, no select chains).  Returns true on a zero / non-finite pivot
// (SimplicialCholesky_impl.h:175-179).  Rows/columns >= nb are identity padding.
__device__ __forceinline__ bool factor_triangle(const double* src, int sq, int sp, double* Tout, int nb) {
  bool bad = false;
  double T[NB][NB], iv[NB];  // entry (q, p) of the input at src[q * sq + p * sp]
  // rectangular loops with compile-time guards: triangular bounds keep the unroller from scalarising T
#pragma unroll
  for (int p = 0; p < NB; p++)
#pragma unroll
    for (int q = 0; q < NB; q++)
      if (q >= p) T[q][p] = (q < nb) ? src[q * sq + p * sp] : (q == p ? 1.0 : 0.0);
#pragma unroll
  for (int p = 0; p < NB; p++) {
    const double d = T[p][p];
    if (p < nb && d == 0.0) bad = true;  // exactly zero only (SimplicialCholesky_impl.h:175-179); NaN / inf flow on like there
    iv[p] = __drcp_rn(d);
#pragma unroll
    for (int q = 0; q < NB; q++) {
      if (q > p) {
        const double lqp = T[q][p] * iv[p];
#pragma unroll
        for (int q2 = 0; q2 < NB; q2++)
          if (q2 >= q) T[q2][q] -= T[q2][p] * lqp;
      }
    }
  }
#pragma unroll
  for (int p = 0; p < NB; p++) {
    Tout[NB * NB + p] = iv[p];
#pragma unroll
    for (int q = 0; q < NB; q++)
      if (q >= p) Tout[q * NB + p] = T[q][p];
  }
  return bad;
}

// ---- second generation of the CTA-per-front kernel -------------------------------------------------
// Phase clocks of the root front of the 10-lap graph (138 pivots; profiles/tools/factor_phase_clocks.py)
// for factor_kernel above: panels 65 % (trailing update 33 %, warp-0 triangle 21 %), fused forward
// solve 15 %, extend-add 14 %.  Changes here, same mathematics and the same panel scheme:
//  * the right-hand side rides along as ROW fs of the front (leading dimension fs + 1): eliminating
//    the augmented matrix [[H, b], [b^T, .]] leaves y (L y = b) in that row under the pivots and the
//    update vector for the parent under the update columns, so the forward solve costs one more row
//    in steps (1b) and (2) instead of a second panel sweep with two barriers per panel;
//  * look-ahead: while warps 1.. run the trailing update of panel k, warp 0 brings the 8 x 8 triangle of
//    panel k+1 up to date, factorises it (a dependent chain of 8 reciprocals, ~1,400 cycles measured in
//    isolation: profiles/tools/tri_lab.cu) and publishes it -- two barriers per panel instead of three
//    and the chain is off the critical path;
//  * extend-add keeps two columns (up to eight loads per lane) in flight (a shared-memory table of the
//    children + four columns in flight was measured too: no change, the phase is not latency-chained);
//  * the trailing update only touches its second 32-row chunk when one exists (fronts of <= 64 rows
//    and the late panels of large fronts have none).
template <bool SMEM>
__global__ void __launch_bounds__(FACTOR_THREADS)
factor2_kernel(SymArgs S, int list_off, const double* __restrict__ V_all, long nV, double* Lv_all, long nL,
               double* Uv_all, long nU, double* Fbig_all, long nFbig, int* status, double* uvec_all, long nUvec,
               double* x_all, int n, int la_idle, int flags) {
  extern __shared__ double smem[];
  // flags & 2: walk the launch list backwards -- the fronts of a class are listed by ascending size, and a level of
  // more than one wave of CTAs ends earlier when its largest fronts start first
  const int early = flags & 1;
  const bool mma_update = SMEM && (flags & 4);  // trailing update by mma.sync.m8n8k4.f64 on 8 x 8 tiles (step (2))
  const int g = S.launch_list[list_off + ((flags & 2) ? gridDim.x - 1 - blockIdx.x : blockIdx.x)];
  const int r = blockIdx.y;
  const int s = S.npiv[g], u = S.nupd[g], fs = s + u;
  const int ld = SMEM ? front_ld(fs) : fs + 1;  // rows 0..fs-1: the front; row fs: the right-hand side; padding
  const double* V = V_all + (size_t)r * nV;
  double* Uv = Uv_all + (size_t)r * nU;
  double* F = SMEM ? smem : (Fbig_all + (size_t)r * nFbig + S.fbig[g]);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  // phase clocks: block 0 of the last launch (the root), or the front SLAM_B200_DBG_FRONT names
  const bool dbgc = S.dbg && blockIdx.y == 0 && tid == 0 && (S.dbg_front >= 0 ? g == S.dbg_front : blockIdx.x == 0);
  if (dbgc) { S.dbg[0] = clock64(); for (int k = 24; k < 32; k++) S.dbg[k] = 0; }
  const bool tlc = S.tl && blockIdx.y == 0 && tid == 0;
  if (tlc) S.tl[TLS * g] = global_ns();
  // the global scratch slab of a front beyond shared memory may still be in use by the previous launch
  if (!SMEM || !early) { pdl_wait_then_release(); if (tlc) S.tl[TLS * g + 1] = global_ns(); }
  for (int t = tid; t < fs * ld; t += nt) F[t] = 0.0;
  __syncthreads();
  if (dbgc) S.dbg[1] = clock64();
  for (int q = S.asm_ptr[g] + tid; q < S.asm_ptr[g + 1]; q += nt) {
    const AsmEntry en = S.asm_entries[q];
    const double* hv = V + en.hoff;
    const int dr = en.meta & 0xff, dc = (en.meta >> 8) & 0xff;
    const bool trans = (en.meta >> 16) & 1, diag = (en.meta >> 17) & 1;
    if (diag) {
      for (int i = 0; i < dr; i++)
        for (int j = 0; j <= i; j++) F[(size_t)(en.c + j) * ld + en.r + i] = __ldcg(hv + i * dc + j);
    } else if (!trans) {
      for (int i = 0; i < dr; i++)
        for (int j = 0; j < dc; j++) F[(size_t)(en.c + j) * ld + en.r + i] = __ldcg(hv + i * dc + j);
    } else {
      for (int i = 0; i < dc; i++)
        for (int j = 0; j < dr; j++) F[(size_t)(en.c + j) * ld + en.r + i] = __ldcg(hv + j * dc + i);
    }
  }
  // ---- the children's row maps, the right-hand side, extend-add ----
  // The row maps (rel) of ALL children go to shared memory at once -- into the panel scratch behind the front, which is
  // idle until the panel loop (>= 21 fs + 448 ints, factor_extra_smem).  They are static, like the rhs of the pivots
  // (V, complete two launches back): with early set both are in place BEFORE the wait, and the lists the rhs gather
  // walks are pulled into L1, so that after the wait only the data of the previous launch is still to be fetched.
  constexpr int ECB = 6, ERC = 3;
  int* srel = reinterpret_cast<int*>(smem + (SMEM ? (size_t)fs * ld : 0));
  const int srel_cap = ((fs + 1) & ~1) + 2 * ((NB + 2) * fs + 8 * NB + 2 * (NB * NB + 2 * NB));
  const int c_first = S.child_ptr[g], c_end = S.child_ptr[g + 1];
  auto stage_maps = [&]() {
    int total = 0;
    for (int ci = c_first; ci < c_end; ci++) {
      const int ch = S.children[ci];
      const int uc = S.nupd[ch];
      if (total + uc <= srel_cap) {
        const int* rel = S.rel + S.rows_ptr[ch];
        for (int i = tid; i < uc; i += nt) srel[total + i] = rel[i];
      }
      total += uc;
    }
    return total;
  };
  const int p0 = S.piv0[g];
  const int* gp = S.gather_ptr + S.frow_ptr[g];
  int rel_total = -1;
  if (SMEM && early) {
    for (int i = tid; i < s; i += nt) F[(size_t)i * ld + fs] = __ldcg(V + S.solver2v[p0 + i]);
    rel_total = stage_maps();
    for (int i = tid; i <= fs; i += nt) prefetch_l1(gp + i);
    {
      const int q0 = gp[0], q1 = gp[fs];
      for (int q = q0 + 32 * tid; q < q1; q += 32 * nt) prefetch_l1(S.gather_src + q);
    }
    for (int ci = c_first + tid; ci < c_end; ci += nt) prefetch_l1(S.uptr + S.children[ci]);
    pdl_wait_then_release();  // everything below reads what the previous launch wrote
    if (tlc) S.tl[TLS * g + 1] = global_ns();
  }
  {  // right-hand side row: rhs of the pivots + the children's update vectors, fixed (child) order
    const double* uvecr = uvec_all + (size_t)r * nUvec;
    for (int i = tid; i < fs; i += nt) {  // thread i wrote the pivot part of entry i itself
      double acc = (SMEM && early) ? F[(size_t)i * ld + fs] : (i < s ? __ldcg(V + S.solver2v[p0 + i]) : 0.0);
      for (int q = gp[i]; q < gp[i + 1]; q++) acc += __ldcg(uvecr + S.gather_src[q]);
      F[(size_t)i * ld + fs] = acc;
    }
  }
  if (rel_total < 0) rel_total = stage_maps();
  const bool staged = rel_total <= srel_cap;  // block-uniform
  __syncthreads();
  if (dbgc) S.dbg[2] = clock64();
  if (tlc) S.tl[TLS * g + 2] = global_ns();
  // A child's Schur complement (uc x uc, lower part valid): warp w takes the columns w, w + nw, ..., up to ECB columns x
  // ERC chunks of 32 rows requested before the first entry is added -- one L2 round trip per child for uc <= 96 on 16
  // warps instead of one per 32 columns (uc = 92: 3 rounds + the row map's before; 4 warps, uc = 47: 2 instead of 6).
  // Children stay sequential with a barrier between them (two children may add to the same entry; the order of the
  // sum is fixed), but the entries of the NEXT child are requested before the current one is added whenever both fit
  // one batch: a front near the leaves has ten and more small children and paid one round trip for each.  (Warp-uniform
  // branches around the slots a small child does not use were measured: slower -- the loads no longer issue back to back.)
  auto load_batch = [&](double (&v)[ECB][ERC], const double* Uc, int uc, int jb, int i0) {
#pragma unroll
    for (int c = 0; c < ECB; c++) {
      const int j = jb + c * nw;
      const double* col = Uc + (size_t)j * uc;
#pragma unroll
      for (int rc = 0; rc < ERC; rc++) {
        const int i = j + i0 + 32 * rc + lane;
        v[c][rc] = (j < uc && i < uc) ? __ldcg(col + i) : 0.0;
      }
    }
  };
  auto add_batch = [&](const double (&v)[ECB][ERC], const int* sr, int uc, int jb, int i0) {
#pragma unroll
    for (int c = 0; c < ECB; c++) {
      const int j = jb + c * nw;
      if (j < uc) {  // warp-uniform
        double* dcol = F + (size_t)sr[j] * ld;
#pragma unroll
        for (int rc = 0; rc < ERC; rc++) {
          const int i = j + i0 + 32 * rc + lane;
          if (i < uc) dcol[sr[i]] += v[c][rc];
        }
      }
    }
  };
  auto one_batch = [&](int uc) { return staged && uc <= nw * ECB && uc <= 32 * ERC; };
  bool all_one = staged;  // every child fits one batch (the rule on SLAM graphs)
  for (int ci = c_first; all_one && ci < c_end; ci++) all_one = one_batch(S.nupd[S.children[ci]]);
  if (all_one && !(flags & 16)) {
    // Two alternating batches: child k is added from one while child k + 1 waits in the other, and the entries of
    // child k + 2 are requested as soon as the batch of child k is free -- nothing waits for a request inside the
    // step that issued it (the single-buffer loop below copied the prefetched batch over after the barrier and so
    // exposed one L2 round trip, ~1 us across the dies, per child: 18 us on the fronts that gather ten children).
    const int nch = c_end - c_first;
    auto child_U = [&](int k) { return Uv + S.uptr[S.children[c_first + k]]; };
    auto child_uc = [&](int k) { return S.nupd[S.children[c_first + k]]; };
    double va[ECB][ERC], vb[ECB][ERC];
    if (nch > 0) load_batch(va, child_U(0), child_uc(0), warp, 0);
    if (nch > 1) load_batch(vb, child_U(1), child_uc(1), warp, 0);
    int roff = 0;
    for (int k = 0; k < nch; k += 2) {
      const int uca = child_uc(k);
      add_batch(va, srel + roff, uca, warp, 0);
      __syncthreads();
      roff += uca;
      if (k + 2 < nch) load_batch(va, child_U(k + 2), child_uc(k + 2), warp, 0);
      if (k + 1 < nch) {
        const int ucb = child_uc(k + 1);
        add_batch(vb, srel + roff, ucb, warp, 0);
        __syncthreads();
        roff += ucb;
        if (k + 3 < nch) load_batch(vb, child_U(k + 3), child_uc(k + 3), warp, 0);
      }
    }
  } else {
    int roff = 0;
    bool have = false;  // va holds the current child already
    double va[ECB][ERC];
    for (int ci = c_first; ci < c_end; ci++) {
      const int ch = S.children[ci];
      const int uc = S.nupd[ch];
      const double* Uc = Uv + S.uptr[ch];
      if (one_batch(uc)) {
        if (!have) load_batch(va, Uc, uc, warp, 0);
        double vb[ECB][ERC];
        bool next = false;
        if (ci + 1 < c_end) {
          const int ch2 = S.children[ci + 1];
          const int uc2 = S.nupd[ch2];
          next = one_batch(uc2);
          if (next) load_batch(vb, Uv + S.uptr[ch2], uc2, warp, 0);
        }
        add_batch(va, srel + roff, uc, warp, 0);
        __syncthreads();
        if (next) {
#pragma unroll
          for (int c = 0; c < ECB; c++)
#pragma unroll
            for (int rc = 0; rc < ERC; rc++) va[c][rc] = vb[c][rc];
        }
        have = next;
      } else {
        const int* sr = srel + roff;
        if (!staged) {  // more children than the scratch holds (not seen on SLAM graphs): one map at a time
          const int* rel = S.rel + S.rows_ptr[ch];
          sr = srel;
          for (int i = tid; i < uc; i += nt) srel[i] = rel[i];
          __syncthreads();
        }
        for (int jb = warp; jb < uc; jb += nw * ECB)
          for (int i0 = 0; i0 < uc - jb; i0 += 32 * ERC) {
            load_batch(va, Uc, uc, jb, i0);
            add_batch(va, sr, uc, jb, i0);
          }
        __syncthreads();
        have = false;
      }
      roff += uc;
    }
  }
  if (dbgc) S.dbg[3] = clock64();
  if (tlc) S.tl[TLS * g + 3] = global_ns();
  double* Sp = reinterpret_cast<double*>(srel + ((fs + 1) & ~1));  // NB x ld: scaled panel (incl. the rhs row)
  double* dinv = Sp + (size_t)NB * ld;                             // 1/d of every pivot (s)
  double* Tsm = dinv + fs;                                          // 2 x (NB x NB triangle + NB reciprocals)
  constexpr int TSZ = NB * NB + NB;
  const int nr = fs + 1;  // rows incl. the right-hand side
  long long dt_ = 0;
  // the first panel's triangle straight from the front; every later one is produced by warp 0 WHILE
  // the other warps run the trailing update of the panel before it (look-ahead, step (2))
  if (warp == 0) {
    if (factor_triangle(F, 1, ld, Tsm, min(NB, s)) && lane == 0) status[2 * r] = 1;
  }
  __syncthreads();
  for (int k0 = 0; k0 < s; k0 += NB) {
    const int nb = min(NB, s - k0);
    double* Pk = F + (size_t)k0 * ld;
    const double* Tc = Tsm + ((k0 / NB) & 1) * TSZ;   // this panel's factorised triangle + reciprocals
    double* Tn = Tsm + (((k0 / NB) & 1) ^ 1) * TSZ;   // the next panel's
    if (dbgc) dt_ = clock64();
    // ---- (1b) one thread per row below the triangle (the rhs row included); the last 64 threads
    // put the triangle and the reciprocals where the later phases read them ----
    {
      const int e = nt - 1 - tid;
      if (e < NB * NB) {
        const int q = e / NB, pp = e % NB;
        if (q >= pp && q < nb) Pk[pp * ld + k0 + q] = Tc[q * NB + pp];
        if (q == 0 && pp < nb) dinv[k0 + pp] = Tc[NB * NB + pp];
      }
    }
    if (k0 + nb + tid < nr) {
      double T[NB][NB], invd[NB];
#pragma unroll
      for (int p = 0; p < NB; p++) {
        invd[p] = Tc[NB * NB + p];
#pragma unroll
        for (int q = p + 1; q < NB; q++) T[q][p] = Tc[q * NB + p];
      }
      for (int i = k0 + nb + tid; i < nr; i += nt) {
        double rr[NB];
#pragma unroll
        for (int p = 0; p < NB; p++) rr[p] = (p < nb) ? Pk[p * ld + i] : 0.0;
#pragma unroll
        for (int p = 0; p < NB; p++) {
          const double rp = rr[p] * invd[p];
          Sp[p * ld + i] = rp;
#pragma unroll
          for (int q = p + 1; q < NB; q++) rr[q] -= rp * T[q][p];
        }
#pragma unroll
        for (int p = 1; p < NB; p++)
          if (p < nb) Pk[p * ld + i] = rr[p];
      }
    }
    if (dbgc) { const long long t_ = clock64(); S.dbg[26] += t_ - dt_; dt_ = t_; }
    __syncthreads();
    if (dbgc) { const long long t_ = clock64(); S.dbg[27] += t_ - dt_; dt_ = t_; }
    // ---- (2) trailing update: columns c0..fs-1, rows j..fs (row fs = rhs); warp 0 instead brings the
    // next panel's triangle up to date (36 entries, one or two per lane), factorises and publishes it ----
    const int c0 = k0 + nb;
    const bool ahead = c0 < s;
    if (ahead && warp == 0) {
      // wait (named barrier 1: this warp + the two producer warps) until the first two column groups
      // -- the next panel's columns -- carry this panel's update, then factorise the triangle from F.
      // (Measured and rejected: producers dropping the triangle into Tn packed for 16-byte loads: +5 %;
      // warps 4/8/12 -- warp 0's scheduler mates -- sitting the step out: the chain drops from 4,500 to
      // 1,700 cycles but the update with 12 instead of 15 warps loses as much.)
      if (mma_update) {
        // tensor-pipe variant: warp 0 brings the tile that holds the next triangle up to date ITSELF (the first task
        // of the update below: two fragments of the panel, two of the scaled panel, two MMAs -- no hand-over from a
        // producer warp, no named barrier) and factorises it where it is, in the accumulator layout, by shuffles
        const int lr = lane >> 2, lk = lane & 3, nb2 = min(NB, s - c0);
        const bool pv0 = lk < nb, pv1 = lk + 4 < nb;
        const double* Pa = Pk + (size_t)lk * ld + c0 + lr;
        const double* Sb = Sp + (size_t)lk * ld + c0 + lr;
        const bool rv = c0 + lr < nr, cv = c0 + lr < fs;
        const double a0 = (rv && pv0) ? -Pa[0] : 0.0, a1 = (rv && pv1) ? -Pa[4 * (size_t)ld] : 0.0;
        const double b0 = (cv && pv0) ? Sb[0] : 0.0, b1 = (cv && pv1) ? Sb[4 * (size_t)ld] : 0.0;
        const int col = c0 + 2 * lk;
        double* cp = F + (size_t)col * ld + c0 + lr;
        const bool v0 = rv && col < fs, v1 = rv && col + 1 < fs;
        double x0 = v0 ? cp[0] : 0.0, x1 = v1 ? cp[ld] : 0.0;
        dmma_upd(x0, x1, a0, b0);
        dmma_upd(x0, x1, a1, b1);
        if (nb2 < NB) {  // the tile reaches beyond the pivots: those entries belong to the trailing matrix
          if (v0) cp[0] = x0;
          if (v1) cp[ld] = x1;
          if (lr >= nb2 || 2 * lk >= nb2) x0 = (lr == 2 * lk) ? 1.0 : 0.0;
          if (lr >= nb2 || 2 * lk + 1 >= nb2) x1 = (lr == 2 * lk + 1) ? 1.0 : 0.0;
        }
        if (dbgc) { const long long t_ = clock64(); S.dbg[25] += t_ - dt_; dt_ = t_; }
        double di0 = 1.0, di1 = 1.0;
        const bool bad = tile_ldlt_unscaled(x0, x1, di0, di1, lr, lk, nb2);
        if (__any_sync(0xffffffffu, bad) && lane == 0) status[2 * r] = 1;
        if (lr >= 2 * lk) Tn[lr * NB + 2 * lk] = x0;
        if (lr >= 2 * lk + 1) Tn[lr * NB + 2 * lk + 1] = x1;
        if (lr == 0) { Tn[NB * NB + 2 * lk] = di0; Tn[NB * NB + 2 * lk + 1] = di1; }
        if (dbgc) { const long long t_ = clock64(); S.dbg[24] += t_ - dt_; dt_ = t_; }
      } else {
      asm volatile("bar.sync 1, 96;" ::: "memory");
      if (dbgc) { const long long t_ = clock64(); S.dbg[25] += t_ - dt_; dt_ = t_; }
      if (factor_triangle(F + (size_t)c0 * ld + c0, 1, ld, Tn, min(NB, s - c0)) && lane == 0) status[2 * r] = 1;
      if (dbgc) { const long long t_ = clock64(); S.dbg[24] += t_ - dt_; dt_ = t_; }
      }
    } else if (!ahead || !(la_idle > 0 && fs - c0 <= la_idle) || (warp & 3) != 0) {
      // la_idle (rows): warps 4, 8, 12 share warp 0's scheduler and FP64 pipe and sit a look-ahead step out when the
      // trailing matrix has at most that many rows (a short update: the triangle chain is what the panel waits for)
      const bool idle_mates = la_idle > 0 && fs - c0 <= la_idle;
      int wi = warp, nwt = nw;
      if (ahead) {
        if (idle_mates) { wi = (warp >> 2) * 3 + (warp & 3) - 1; nwt = (nw >> 2) * 3; }
        else { wi = warp - 1; nwt = nw - 1; }
      }
      if (mma_update) {
        // Trailing update on the fp64 tensor pipe, operands and accumulators straight from the front in shared memory.
        // The scalar update below moves 96 shared-memory wavefronts per 64 x 4 x 8 block of multiply-adds (A 32, the
        // broadcast B values 32, C in and out 32) and the ONE shared-memory pipe of the SM is what the panel step waits
        // for (phase clocks of a 46-pivot / 138-row front: update and look-ahead triangle both ~5,400 cycles per panel,
        // the triangle as slow with its scheduler mates idle).  As 8 x 8 tiles with the two column tiles of a pair
        // sharing the A fragments the same work moves 28 wavefronts per 16 x 8 x 8 block: 56 for the work of 96.
        // Tasks = (pair of column tiles, row tile) below or on the diagonal, dealt round-robin to the warps; task 0
        // holds the next panel's triangle and is warp 0's own when it looks ahead (above).  Tiles on the diagonal are
        // written whole: the strict upper triangle of a front is never read.
        const int Tn = (nr - c0 + 7) >> 3, Tc = (fs - c0 + 7) >> 3;  // Tn >= Tc: the rhs row is one more row
        const int lr = lane >> 2, lk = lane & 3;
        // operand addresses of this lane relative to (row tile, column tile); k-halves beyond the panel's pivots read 0
        const bool pv0 = lk < nb, pv1 = lk + 4 < nb;
        const double* Pa0 = Pk + (size_t)lk * ld + c0 + lr;        // A, first k-half: panel column lk, row c0 + 8 I + lr
        const double* Pa1 = Pa0 + 4 * (size_t)ld;
        const double* Sb0 = Sp + (size_t)lk * ld + c0 + lr;        // B: scaled panel column lk, "row" = column c0 + 8 J + lr
        const double* Sb1 = Sb0 + 4 * (size_t)ld;
        double* Fc = F + (size_t)(c0 + 2 * lk) * ld + c0 + lr;     // C: column c0 + 8 J + 2 lk (+ 1), row c0 + 8 I + lr
        // flat order of the tasks: Jp = 0: I = 0 .. Tn-1; Jp = 1: I = 2 .. Tn-1; ...  (position 0 = the next triangle's
        // tile: warp 0's when it looks ahead).  A warp takes a CONTIGUOUS range of positions: consecutive row tiles of one
        // pair of column tiles share the B fragments, and two row tiles go through loads -> MMAs -> stores together.
        const int nJ = (Tc + 1) >> 1;
        const int total = nJ * Tn - nJ * (nJ - 1);  // sum over Jp of (Tn - 2 Jp)
        const int first = ahead ? 1 : 0;
        const int chunk = (total - first + nwt - 1) / nwt;
        int pos = first + wi * chunk;
        const int pend = min(pos + chunk, total);
        int Jp = 0, I = 0;
        for (int n = pos; Jp < nJ; Jp++) {
          const int left = Tn - 2 * Jp;
          if (n < left) { I = 2 * Jp + n; break; }
          n -= left;
        }
        while (pos < pend) {
          const int len = min(pend - pos, Tn - I);
          const int colr = 16 * Jp;  // relative to c0 (+ lr: in Sb0 / Sb1)
          const bool cv0 = c0 + colr + lr < fs, cv1 = c0 + colr + lr + 8 < fs;
          const double b00 = (cv0 && pv0) ? Sb0[colr] : 0.0, b01 = (cv0 && pv1) ? Sb1[colr] : 0.0;
          const double b10 = (cv1 && pv0) ? Sb0[colr + 8] : 0.0, b11 = (cv1 && pv1) ? Sb1[colr + 8] : 0.0;
          const bool two = 2 * Jp + 1 < Tc;
          const int col0 = c0 + colr + 2 * lk;
          const bool cA0 = col0 < fs, cA1 = col0 + 1 < fs, cB0 = two && col0 + 8 < fs, cB1 = two && col0 + 9 < fs;
          for (int q = 0; q < len; q += 2) {
            const int Ia = I + q;
            const int ra = 8 * Ia, rb = ra + 8;
            const bool ha = c0 + ra + lr < nr, hb = q + 1 < len && c0 + rb + lr < nr;
            const bool sa = Ia > 2 * Jp;  // row tile Ia reaches the second column tile (Ia + 1 always does)
            double* cpa = Fc + (size_t)colr * ld + ra;
            double* cpb = cpa + 8;
            double* cqa = cpa + 8 * (size_t)ld;
            double* cqb = cqa + 8;
            const bool va0 = ha && cA0, va1 = ha && cA1, vb0 = hb && cA0, vb1 = hb && cA1;
            const bool wa0 = ha && sa && cB0, wa1 = ha && sa && cB1, wb0 = hb && cB0, wb1 = hb && cB1;
            const double aa0 = (ha && pv0) ? -Pa0[ra] : 0.0, aa1 = (ha && pv1) ? -Pa1[ra] : 0.0;
            const double ab0 = (hb && pv0) ? -Pa0[rb] : 0.0, ab1 = (hb && pv1) ? -Pa1[rb] : 0.0;
            double xa0 = va0 ? cpa[0] : 0.0, xa1 = va1 ? cpa[ld] : 0.0, xb0 = vb0 ? cpb[0] : 0.0, xb1 = vb1 ? cpb[ld] : 0.0;
            double ya0 = wa0 ? cqa[0] : 0.0, ya1 = wa1 ? cqa[ld] : 0.0, yb0 = wb0 ? cqb[0] : 0.0, yb1 = wb1 ? cqb[ld] : 0.0;
            dmma_upd(xa0, xa1, aa0, b00);
            dmma_upd(xb0, xb1, ab0, b00);
            dmma_upd(ya0, ya1, aa0, b10);
            dmma_upd(yb0, yb1, ab0, b10);
            dmma_upd(xa0, xa1, aa1, b01);
            dmma_upd(xb0, xb1, ab1, b01);
            dmma_upd(ya0, ya1, aa1, b11);
            dmma_upd(yb0, yb1, ab1, b11);
            if (va0) cpa[0] = xa0;
            if (va1) cpa[ld] = xa1;
            if (vb0) cpb[0] = xb0;
            if (vb1) cpb[ld] = xb1;
            if (wa0) cqa[0] = ya0;
            if (wa1) cqa[ld] = ya1;
            if (wb0) cqb[0] = yb0;
            if (wb1) cqb[ld] = yb1;
          }
          pos += len;
          Jp++;
          I = 2 * Jp;
        }
      } else {
      bool owe = ahead && wi < 2;  // producer of the next panel's columns: signal warp 0 after the first group
      for (int jg = c0 + 4 * wi; jg < fs; jg += 4 * nwt) {
        double B[4][NB];
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int p = 0; p < NB; p++) B[b][p] = (jg + b < fs) ? Sp[p * ld + jg + b] : 0.0;
        double* Cj = F + (size_t)jg * ld;
        int i = jg + lane;
        for (; i - lane + 32 < nr; i += 64) {  // warp-uniform: a second 32-row chunk exists
          const int i2 = i + 32;
          const bool v2 = i2 < nr;
          double A[NB], A2[NB];
#pragma unroll
          for (int p = 0; p < NB; p++) {
            A[p] = (p < nb) ? Pk[p * ld + i] : 0.0;
            A2[p] = (p < nb && v2) ? Pk[p * ld + i2] : 0.0;
          }
#pragma unroll
          for (int b = 0; b < 4; b++) {
            if (jg + b < fs) {
              const bool w1 = i >= jg + b;
              double acc = w1 ? Cj[b * ld + i] : 0.0;
              double acc2 = v2 ? Cj[b * ld + i2] : 0.0;
#pragma unroll
              for (int p = 0; p < NB; p++) {
                acc -= A[p] * B[b][p];
                acc2 -= A2[p] * B[b][p];
              }
              if (w1) Cj[b * ld + i] = acc;
              if (v2) Cj[b * ld + i2] = acc2;
            }
          }
        }
        if (i - lane < nr) {  // last, single chunk
          const bool v1 = i < nr;
          double A[NB];
#pragma unroll
          for (int p = 0; p < NB; p++) A[p] = (p < nb && v1) ? Pk[p * ld + i] : 0.0;
#pragma unroll
          for (int b = 0; b < 4; b++) {
            if (jg + b < fs) {
              const bool w1 = v1 && i >= jg + b;
              double acc = w1 ? Cj[b * ld + i] : 0.0;
#pragma unroll
              for (int p = 0; p < NB; p++) acc -= A[p] * B[b][p];
              if (w1) Cj[b * ld + i] = acc;
            }
          }
        }
        if (owe) { __threadfence_block(); asm volatile("bar.arrive 1, 96;" ::: "memory"); owe = false; }
      }
      if (owe) { __threadfence_block(); asm volatile("bar.arrive 1, 96;" ::: "memory"); }
      }
    }
    if (dbgc) { const long long t_ = clock64(); S.dbg[28] += t_ - dt_; dt_ = t_; }
    __syncthreads();
    if (dbgc) { const long long t_ = clock64(); S.dbg[29] += t_ - dt_; dt_ = t_; }
  }
  if (dbgc) { S.dbg[4] = clock64(); S.dbg[5] = S.dbg[4]; }
  if (tlc) S.tl[TLS * g + 4] = global_ns();
  // forward-solve results out of the rhs row: z = D^-1 y under the pivots, update vector for the parent
  {
    const int p0 = S.piv0[g];
    double* xr = x_all + (size_t)r * n;
    double* uo = uvec_all + (size_t)r * nUvec + S.rows_ptr[g];
    for (int i = tid; i < fs; i += nt) {
      const double v = F[(size_t)i * ld + fs];
      if (i < s) xr[p0 + i] = v * dinv[i];
      else uo[i - s] = v;
    }
  }
  double* Lg = Lv_all + (size_t)r * nL + S.lptr[g];
  for (int j = warp; j < s; j += nw) {
    const double* col = F + (size_t)j * ld;
    const double d = col[j];
    const double inv = dinv[j];
    double* out = Lg + (size_t)j * fs;
    for (int i = lane; i < fs; i += 32) out[i] = i < j ? 0.0 : (i == j ? d : col[i] * inv);
  }
  double* Ug = Uv + S.uptr[g];
  for (int j = warp; j < u; j += nw) {
    const double* col = F + (size_t)(s + j) * ld + s;
    double* out = Ug + (size_t)j * u;
    for (int i = j + lane; i < u; i += 32) out[i] = col[i];
  }
  if ((flags & 8) && u == 0) {
    // A root (no update rows, no parent): its backward solve L^T x = z needs nothing but this front, which is still in
    // shared memory -- done here, the level's backward launch (stage 150 KB of L again, 18 dependent panel steps in a
    // CTA of its own: 22 us on the 10-lap graph) is not enqueued.  Same arithmetic, same order as backward_kernel:
    // L[i][k] = F[i, k] / d_k exactly as it is written out above.
    __syncthreads();  // the panel scratch (Sp) is free again; F's columns are only read from here on
    double* xs = Sp;        // s
    double* xo = Sp + ld;   // NB: the panel being solved
    for (int i = tid; i < s; i += nt) xs[i] = F[(size_t)i * ld + fs] * dinv[i];
    __syncthreads();
    for (int pan = (s + NB - 1) / NB - 1; pan >= 0; pan--) {
      const int k0 = pan * NB;
      const int nb = min(NB, s - k0);
      if (warp == 0) {
        double xp[NB];
#pragma unroll
        for (int p = 0; p < NB; p++) xp[p] = (p < nb) ? xs[k0 + p] : 0.0;
#pragma unroll
        for (int p = NB - 1; p >= 0; p--)
#pragma unroll
          for (int q = 0; q < p; q++)
            if (p < nb) xp[q] -= (F[(size_t)(k0 + q) * ld + k0 + p] * dinv[k0 + q]) * xp[p];
        if (lane == 0) {
#pragma unroll
          for (int p = 0; p < NB; p++)
            if (p < nb) { xo[p] = xp[p]; xs[k0 + p] = xp[p]; }
        }
      }
      __syncthreads();
      for (int k = tid; k < k0; k += nt) {
        const double* col = F + (size_t)k * ld + k0;
        const double inv = dinv[k];
        double acc = xs[k];
#pragma unroll
        for (int p = 0; p < NB; p++)
          if (p < nb) acc -= (col[p] * inv) * xo[p];
        xs[k] = acc;
      }
      __syncthreads();
    }
    double* xr = x_all + (size_t)r * n + S.piv0[g];
    for (int i = tid; i < s; i += nt) xr[i] = xs[i];
  }
  __syncthreads();
  if (dbgc) { S.dbg[6] = clock64(); S.dbg[7] = s; S.dbg[8] = fs; S.dbg[9] = S.child_ptr[g + 1] - S.child_ptr[g]; }
  if (tlc) S.tl[TLS * g + 5] = global_ns();
}

// ================================================================================================
// Warp-per-front kernels for small fronts (fs <= 64 rows): no block barriers, no redundant work.
// Lane l owns front rows l and l + 32.  Most fronts of a trackdrive graph are this small (the
// leaves of the assembly tree) and all fronts of the batched Monte-Carlo configuration are, so
// these kernels carry the throughput; the CTA-per-front kernels above carry the few large fronts
// near the root.
// ================================================================================================
constexpr int TINY_WARPS = 4;   // fronts per CTA
constexpr int TNB = 4;          // pivots per panel

// The front is stored PACKED (lower triangle only, column j holds rows j..fs-1 at offset
// j*fs - j(j-1)/2): half the shared memory of a square front = twice the fronts resident per SM,
// which is what bounds this latency-dominated kernel.
__device__ __forceinline__ int tri_off(int j, int fs) { return j * fs - ((j * (j - 1)) >> 1) - j; }  // + row

__global__ void __launch_bounds__(TINY_WARPS * 32)
factor_tiny_kernel(SymArgs S, int list_off, int count, int slab, const double* __restrict__ V_all, long nV,
                   double* Lv_all, long nL, double* Uv_all, long nU, int* status, double* uvec_all, long nUvec,
                   double* x_all, int n) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int fi = blockIdx.x * TINY_WARPS + wid;
  if (fi >= count) return;  // no block-wide barrier below
  const int g = S.launch_list[list_off + fi];
  const int r = blockIdx.y;
  const int s = S.npiv[g], u = S.nupd[g], fs = s + u;
  const double* V = V_all + (size_t)r * nV;
  double* Uv = Uv_all + (size_t)r * nU;
  double* F = smem + (size_t)wid * slab;  // packed lower triangle: F[tri_off(j, fs) + i], i >= j
  const int ntri = (fs * (fs + 1)) >> 1;
  long long tclk = S.dbg ? clock64() : 0;
  for (int t = lane; t < ntri; t += 32) F[t] = 0.0;
  __syncwarp();
  if (S.dbg && lane == 0) { const long long t_ = clock64(); atomicAdd((unsigned long long*)&S.dbg[16 + 0], (unsigned long long)(t_ - tclk)); tclk = t_; }
  for (int q = S.asm_ptr[g] + lane; q < S.asm_ptr[g + 1]; q += 32) {
    const AsmEntry en = S.asm_entries[q];
    const double* hv = V + en.hoff;
    const int dr = en.meta & 0xff, dc = (en.meta >> 8) & 0xff;
    const bool trans = (en.meta >> 16) & 1, diag = (en.meta >> 17) & 1;
    if (diag) {
      for (int i = 0; i < dr; i++)
        for (int j = 0; j <= i; j++) F[tri_off(en.c + j, fs) + en.r + i] = hv[i * dc + j];
    } else if (!trans) {
      for (int i = 0; i < dr; i++)
        for (int j = 0; j < dc; j++) F[tri_off(en.c + j, fs) + en.r + i] = hv[i * dc + j];
    } else {
      for (int i = 0; i < dc; i++)
        for (int j = 0; j < dr; j++) F[tri_off(en.c + j, fs) + en.r + i] = hv[j * dc + i];
    }
  }
  __syncwarp();
  if (S.dbg && lane == 0) { const long long t_ = clock64(); atomicAdd((unsigned long long*)&S.dbg[16 + 1], (unsigned long long)(t_ - tclk)); tclk = t_; }
  // extend-add the children's Schur complements (children of a small front are small: uc <= 64)
  for (int ci = S.child_ptr[g]; ci < S.child_ptr[g + 1]; ci++) {
    const int ch = S.children[ci];
    const int uc = S.nupd[ch];
    const double* Uc = Uv + S.uptr[ch];
    const int* rel = S.rel + S.rows_ptr[ch];
    const int i0 = lane, i1 = lane + 32;
    const int rl0 = i0 < uc ? rel[i0] : 0, rl1 = i1 < uc ? rel[i1] : 0;
    // four columns per step, all eight global loads issued before the first shared-memory update:
    // the loop is bound by the latency of these loads, not by their volume
    for (int j = 0; j < uc; j += 4) {
      double v0[4], v1[4];
      int relj[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int jj = j + q;
        const int ra = __shfl_sync(0xffffffffu, rl0, jj & 31), rb = __shfl_sync(0xffffffffu, rl1, jj & 31);
        relj[q] = jj < 32 ? ra : rb;
        const double* col = Uc + (size_t)jj * uc;
        v0[q] = (jj < uc && i0 >= jj && i0 < uc) ? __ldg(col + i0) : 0.0;
        v1[q] = (jj < uc && i1 >= jj && i1 < uc) ? __ldg(col + i1) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int jj = j + q;
        double* dst = F + tri_off(relj[q], fs);
        if (jj < uc && i0 >= jj && i0 < uc) dst[rl0] += v0[q];
        if (jj < uc && i1 >= jj && i1 < uc) dst[rl1] += v1[q];
      }
    }
    __syncwarp();
  }
  if (S.dbg && lane == 0) { const long long t_ = clock64(); atomicAdd((unsigned long long*)&S.dbg[16 + 2], (unsigned long long)(t_ - tclk)); tclk = t_; }
  // right-looking LDL^T, panels of TNB pivots held in registers; the panel's triangle and the
  // column factors travel between lanes by shuffle
  bool bad = false;
  for (int k0 = 0; k0 < s; k0 += TNB) {
    const int nb = min(TNB, s - k0);
    const int r0 = k0 + lane, r1 = r0 + 32;
    double a0[TNB], a1[TNB], inv[TNB];
#pragma unroll
    for (int p = 0; p < TNB; p++) {
      // lanes above the diagonal of the panel (r0 < k0 + p) hold 0: the packed layout has no such entry
      a0[p] = (p < nb) ? ((r0 < fs && r0 >= k0 + p) ? F[tri_off(k0 + p, fs) + r0] : 0.0) : (lane == p ? 1.0 : 0.0);
      a1[p] = (p < nb && r1 < fs) ? F[tri_off(k0 + p, fs) + r1] : 0.0;
    }
#pragma unroll
    for (int p = 0; p < TNB; p++) {
      const double d = __shfl_sync(0xffffffffu, a0[p], p);
      if (p < nb && d == 0.0) bad = true;  // SimplicialCholesky_impl.h:175-179: an exactly zero pivot fails, NaN / inf flow on
      inv[p] = __drcp_rn(d);
#pragma unroll
      for (int q = p + 1; q < TNB; q++) {
        const double f = __shfl_sync(0xffffffffu, a0[p], q) * inv[p];
        a0[q] -= a0[p] * f;
        a1[q] -= a1[p] * f;
      }
    }
    // the panel's columns go back SCALED (l_ip = a_ip / d_p, the entries of L; d_p stays on the
    // diagonal): the trailing update below needs a_ip (in registers) times l_jp, and reads l_jp of its
    // column j from here with broadcast loads -- no shuffles, no scratch; shuffles and shared-memory
    // accesses share one pipe, and the write of L at the end needs no division any more
#pragma unroll
    for (int p = 0; p < TNB; p++) {
      if (p < nb) {
        if (r0 < fs && r0 >= k0 + p) F[tri_off(k0 + p, fs) + r0] = (r0 == k0 + p) ? a0[p] : a0[p] * inv[p];
        if (r1 < fs) F[tri_off(k0 + p, fs) + r1] = a1[p] * inv[p];
      }
    }
    __syncwarp();
    const double* pc[TNB];  // the panel's columns (row index added on use)
#pragma unroll
    for (int p = 0; p < TNB; p++) {
      pc[p] = F + tri_off(k0 + min(p, nb - 1), fs);
      if (p >= nb) { a0[p] = 0.0; a1[p] = 0.0; }  // padding pivots of a short last panel contribute nothing
    }
    int j = k0 + nb;
    for (; j + 1 < fs; j += 2) {  // two independent columns per step
      double* cA = F + tri_off(j, fs);
      double* cB = F + tri_off(j + 1, fs);
      const bool wA0 = r0 >= j && r0 < fs, wB0 = r0 >= j + 1 && r0 < fs;
      const bool wA1 = r1 >= j && r1 < fs, wB1 = r1 >= j + 1 && r1 < fs;
      double aA0 = wA0 ? cA[r0] : 0.0, aA1 = wA1 ? cA[r1] : 0.0;
      double aB0 = wB0 ? cB[r0] : 0.0, aB1 = wB1 ? cB[r1] : 0.0;
      double fa[TNB], fb[TNB];
#pragma unroll
      for (int p = 0; p < TNB; p++) {  // broadcast reads; pivots beyond nb contribute a_ip = 0
        fa[p] = pc[p][j];
        fb[p] = pc[p][j + 1];
      }
#pragma unroll
      for (int p = 0; p < TNB; p++) {
        aA0 -= a0[p] * fa[p]; aA1 -= a1[p] * fa[p];
        aB0 -= a0[p] * fb[p]; aB1 -= a1[p] * fb[p];
      }
      if (wA0) cA[r0] = aA0;
      if (wA1) cA[r1] = aA1;
      if (wB0) cB[r0] = aB0;
      if (wB1) cB[r1] = aB1;
    }
    for (; j < fs; j++) {
      double* cA = F + tri_off(j, fs);
      const bool wA0 = r0 >= j && r0 < fs, wA1 = r1 >= j && r1 < fs;
      double acc0 = wA0 ? cA[r0] : 0.0, acc1 = wA1 ? cA[r1] : 0.0;
#pragma unroll
      for (int p = 0; p < TNB; p++) {
        const double cj = pc[p][j];
        acc0 -= a0[p] * cj;
        acc1 -= a1[p] * cj;
      }
      if (wA0) cA[r0] = acc0;
      if (wA1) cA[r1] = acc1;
    }
    __syncwarp();
  }
  if (bad && lane == 0) status[2 * r] = 1;
  if (S.dbg && lane == 0) { const long long t_ = clock64(); atomicAdd((unsigned long long*)&S.dbg[16 + 3], (unsigned long long)(t_ - tclk)); tclk = t_; }
  // ---- fused forward solve (see factor_kernel): lane l owns rows l and l + 32 of w ----
  {
    const int p0 = S.piv0[g];
    const double* uvecr = uvec_all + (size_t)r * nUvec;
    const int* gp = S.gather_ptr + S.frow_ptr[g];
    const int i0 = lane, i1 = lane + 32;
    double w0 = 0.0, w1 = 0.0;
    if (i0 < fs) {
      w0 = i0 < s ? V[S.solver2v[p0 + i0]] : 0.0;
      for (int q = gp[i0]; q < gp[i0 + 1]; q++) w0 += uvecr[S.gather_src[q]];
    }
    if (i1 < fs) {
      w1 = i1 < s ? V[S.solver2v[p0 + i1]] : 0.0;
      for (int q = gp[i1]; q < gp[i1 + 1]; q++) w1 += uvecr[S.gather_src[q]];
    }
    for (int k = 0; k < s; k++) {
      const double* col = F + tri_off(k, fs);  // l_ik below the diagonal, d_k on it
      const double dk = col[k];
      const double wa = __shfl_sync(0xffffffffu, w0, k & 31), wb = __shfl_sync(0xffffffffu, w1, k & 31);
      const double yk = k < 32 ? wa : wb;
      if (i0 > k && i0 < fs) w0 -= col[i0] * yk;
      if (i1 > k && i1 < fs) w1 -= col[i1] * yk;
      if (lane == (k & 31)) { if (k < 32) w0 = yk * __drcp_rn(dk); else w1 = yk * __drcp_rn(dk); }  // z_k = y_k / d_k
    }
    double* xr = x_all + (size_t)r * n;
    double* uo = uvec_all + (size_t)r * nUvec + S.rows_ptr[g];
    if (i0 < s) xr[p0 + i0] = w0; else if (i0 < fs) uo[i0 - s] = w0;
    if (i1 < s) xr[p0 + i1] = w1; else if (i1 < fs) uo[i1 - s] = w1;
  }
  if (S.dbg && lane == 0) { const long long t_ = clock64(); atomicAdd((unsigned long long*)&S.dbg[16 + 4], (unsigned long long)(t_ - tclk)); tclk = t_; }
  double* Lg = Lv_all + (size_t)r * nL + S.lptr[g];
  for (int j = 0; j < s; j++) {
    const double* col = F + tri_off(j, fs);
    double* out = Lg + (size_t)j * fs;
    for (int i = lane; i < fs; i += 32) out[i] = i < j ? 0.0 : col[i];
  }
  double* Ug = Uv + S.uptr[g];
  for (int j = 0; j < u; j++) {
    const double* col = F + tri_off(s + j, fs) + s;
    double* out = Ug + (size_t)j * u;
    for (int i = j + lane; i < u; i += 32) out[i] = col[i];
  }
  if (S.dbg && lane == 0) { const long long t_ = clock64(); atomicAdd((unsigned long long*)&S.dbg[16 + 5], (unsigned long long)(t_ - tclk)); atomicAdd((unsigned long long*)&S.dbg[16 + 6], 1ull); }
}

__global__ void __launch_bounds__(TINY_WARPS * 32)
forward_tiny_kernel(SymArgs S, int list_off, int count, const double* __restrict__ V_all, long nV,
                    const double* __restrict__ Lv_all, long nL, double* uvec_all, long nUvec, double* x_all, int n) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int fi = blockIdx.x * TINY_WARPS + wid;
  if (fi >= count) return;
  const int g = S.launch_list[list_off + fi];
  const int r = blockIdx.y;
  const int s = S.npiv[g], u = S.nupd[g], fs = s + u, p0 = S.piv0[g];
  const double* V = V_all + (size_t)r * nV;
  const double* Lg = Lv_all + (size_t)r * nL + S.lptr[g];
  double* uvec = uvec_all + (size_t)r * nUvec;
  double* x = x_all + (size_t)r * n;
  const int* gp = S.gather_ptr + S.frow_ptr[g];
  const int i0 = lane, i1 = lane + 32;
  double w0 = 0.0, w1 = 0.0;
  if (i0 < fs) {
    w0 = i0 < s ? V[S.solver2v[p0 + i0]] : 0.0;
    for (int q = gp[i0]; q < gp[i0 + 1]; q++) w0 += uvec[S.gather_src[q]];
  }
  if (i1 < fs) {
    w1 = i1 < s ? V[S.solver2v[p0 + i1]] : 0.0;
    for (int q = gp[i1]; q < gp[i1 + 1]; q++) w1 += uvec[S.gather_src[q]];
  }
  for (int k0 = 0; k0 < s; k0 += 4) {  // four columns of L in flight, then the dependent updates
    double l0[4], l1[4];
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const int k = k0 + p;
      l0[p] = (k < s && i0 > k && i0 < fs) ? Lg[(size_t)k * fs + i0] : 0.0;
      l1[p] = (k < s && i1 > k && i1 < fs) ? Lg[(size_t)k * fs + i1] : 0.0;
    }
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const int k = k0 + p;
      const double wa = __shfl_sync(0xffffffffu, w0, k & 31), wb = __shfl_sync(0xffffffffu, w1, k & 31);
      const double wk = k < 32 ? wa : wb;
      w0 -= l0[p] * wk;
      w1 -= l1[p] * wk;
    }
  }
  if (i0 < s) x[p0 + i0] = w0 / Lg[(size_t)i0 * fs + i0];
  else if (i0 < fs) uvec[S.rows_ptr[g] + i0 - s] = w0;
  if (i1 < s) x[p0 + i1] = w1 / Lg[(size_t)i1 * fs + i1];
  else if (i1 < fs) uvec[S.rows_ptr[g] + i1 - s] = w1;
}

__global__ void __launch_bounds__(TINY_WARPS * 32)
backward_tiny_kernel(SymArgs S, int list_off, int count, const double* __restrict__ Lv_all, long nL,
                     double* x_all, int n) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int fi = blockIdx.x * TINY_WARPS + wid;
  if (fi >= count) return;
  const int g = S.launch_list[list_off + fi];
  const int r = blockIdx.y;
  const int s = S.npiv[g], u = S.nupd[g], fs = s + u, p0 = S.piv0[g];
  const double* Lg = Lv_all + (size_t)r * nL + S.lptr[g];
  double* x = x_all + (size_t)r * n;
  const int* rows = S.upd_rows + S.rows_ptr[g];
  const int i0 = lane, i1 = lane + 32;
  const int r0 = (i0 >= s && i0 < fs) ? rows[i0 - s] : 0, r1 = (i1 >= s && i1 < fs) ? rows[i1 - s] : 0;
  const bool tlc = S.tl && r == 0 && lane == 0;
  if (tlc) S.tl[TLS * g + 6] = global_ns();
  pdl_wait_then_release();  // the parents' x comes from the previous launch
  if (tlc) S.tl[TLS * g + 7] = global_ns();
  double x0 = 0.0, x1 = 0.0;
  if (i0 < fs) x0 = __ldcg(i0 < s ? x + p0 + i0 : x + r0);
  if (i1 < fs) x1 = __ldcg(i1 < s ? x + p0 + i1 : x + r1);
  // The four columns of step kb - 4 are requested before the columns of step kb are used: L comes from L2 (written by
  // the factor kernels, often on the other die) and a front of 40 pivots would otherwise pay ten round trips in a row.
  auto load_cols = [&](int kb, double (&l0)[4], double (&l1)[4]) {
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const int k = kb + p;
      l0[p] = (kb >= 0 && k < s && i0 > k && i0 < fs) ? Lg[(size_t)k * fs + i0] : 0.0;
      l1[p] = (kb >= 0 && k < s && i1 > k && i1 < fs) ? Lg[(size_t)k * fs + i1] : 0.0;
    }
  };
  double n0[4], n1[4];
  load_cols(((s - 1) / 4) * 4, n0, n1);
  for (int kb = ((s - 1) / 4) * 4; kb >= 0; kb -= 4) {
    double l0[4], l1[4];
#pragma unroll
    for (int p = 0; p < 4; p++) { l0[p] = n0[p]; l1[p] = n1[p]; }
    load_cols(kb - 4, n0, n1);
#pragma unroll
    for (int p = 3; p >= 0; p--) {
      const int k = kb + p;
      if (k < s) {  // warp-uniform
        double t = l0[p] * x0 + l1[p] * x1;
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (k < 32) { if (lane == k) x0 -= t; }
        else if (lane == k - 32) x1 -= t;
      }
    }
  }
  if (i0 < s) x[p0 + i0] = x0;
  if (i1 < s) x[p0 + i1] = x1;
  if (tlc) S.tl[TLS * g + 8] = global_ns();
}

// ================================================================================================
// TILED kernels for replica batches (tileplan.h): one warp per (front, replica); the front is kept as
// 8 x 8 fp64 tiles in shared memory, the accumulator shape of mma.sync.m8n8k4.f64, so a tile update
// C -= X L^T is two tensor-pipe instructions instead of 64 predicated DFMA + their loads.  (The fp64 MMA
// runs at the rate of the DFMA pipe on B200 -- profiles/r02_dmma_lab.log: 36.7 TFLOP/s either way -- the
// gain is the instruction count: ncu counted 12,000 warp instructions per front for factor_tiny_kernel,
// 350 of them useful FMAs.)  The four warps of a CTA work on the SAME front of four consecutive
// replicas: the assembly list is read once from L2 and three times from L1.
//
// Per front (local layout and storage: tileplan.h):
//   zero the tile triangle, 1.0 on the padding pivots; V items are stored, children's items added;
//   for every pivot tile column K:
//     (a) every lane factorises the 8 x 8 diagonal tile redundantly in registers (LDL^T, unscaled
//         right-looking, SimplicialCholesky_impl.h:122-191 restricted to the tile), lane j < 8 then
//         solves column j of inv(L11) (28 FMA, lane-independent code); W = inv(L11)^T goes to scratch;
//     (b) for every tile row I below: X = A_IK W (2 MMA; = L_IK D), L_IK = X D^-1 stored in place,
//         -X to scratch (accumulator layout -> A-operand layout), then C_IJ += (-X) L_JK^T for
//         K < J <= I (2 MMA per tile).
//   The rhs row is just a row of the last tile row: it leaves as z = D^-1 L^-1 b under the pivots and as
//   the update vector under the update columns.  Finally the whole triangle is copied to HBM (unit
//   stride, 16-byte stores) and z goes to x.
// ================================================================================================
constexpr int TILE_SCRATCH = 0;  // nothing behind the tile triangle any more

struct TileArgs {
  const int *npiv, *nupd, *piv0, *rows_ptr, *upd_rows, *list, *item_ptr, *item_nv, *child_ptr, *children;
  const long* fptr;
  const int2* items;
};

__device__ __forceinline__ void dmma_acc(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ int tile_base(int I, int J) { return (((I * (I + 1)) >> 1) + J) << 6; }
// element (i, j) of the tile triangle; rows 2, 3, 6, 7 of a tile keep their 4-column halves swapped (tileplan.h)
__device__ __forceinline__ int tile_in(int gi, int cj) { return (gi << 3) + (cj ^ ((gi & 2) << 1)); }
__device__ __forceinline__ int tile_at(int i, int j) { return tile_base(i >> 3, j >> 3) + tile_in(i & 7, j & 7); }

// 1/d to within an ulp or two: hardware seed (about 20 bits) + two Newton steps.  __drcp_rn is correctly rounded
// but costs 35 instructions (ncu: 19 % of the kernel's instructions sat on that line); the result is compared with
// the CPU at 1e-6, not bit for bit.
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  return fma(x, e, x);
}
// SimplicialCholesky_impl.h:175-179 fails on an exactly zero pivot and on nothing else: a NaN or infinite pivot flows
// on and poisons everything connected to it, which is what the reference's map looks like after an optimise over a
// graph that holds a NaN cone (an absent objectId: azimuth 0).  A denormal pivot counts as zero here: the seeded
// reciprocal flushes it.
__device__ __forceinline__ bool bad_pivot(double d) {
  return ((__double2hiint(d) >> 20) & 0x7ff) == 0;
}

// LDL^T of one 8 x 8 tile held in the accumulator layout of the warp (lane (g, t): entries (g, 2t), (g, 2t+1)),
// in place and WITHOUT redundancy: step p fetches the four entries of column p a lane needs (its row, the pivot,
// the rows of its two columns) by shuffle.  Afterwards the lane holds the unit-lower L (d on the diagonal, whatever
// the symmetric updates left above it) and 1/d of its two columns.  (The first version factorised the tile
// redundantly in the registers of every lane: two thirds of the kernel's instructions.)
__device__ __forceinline__ bool tile_ldlt(double& a0, double& a1, double& di0, double& di1, int g, int t) {
  bool bad = false;
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int hp = p >> 1;
    const double v = (p & 1) ? a1 : a0;
    const double rgp = __shfl_sync(0xffffffffu, v, g * 4 + hp);      // T[g][p]
    const double dp = __shfl_sync(0xffffffffu, v, p * 4 + hp);       // T[p][p]
    const double c0 = __shfl_sync(0xffffffffu, v, 8 * t + hp);       // T[2t][p]
    const double c1 = __shfl_sync(0xffffffffu, v, 8 * t + 4 + hp);   // T[2t+1][p]
    bad |= bad_pivot(dp);
    const double inv = fast_rcp(dp);
    const double lgp = rgp * inv;
    if (2 * t > p) a0 -= lgp * c0;
    if (2 * t + 1 > p) a1 -= lgp * c1;
    if (t == hp) {
      if (p & 1) { di1 = inv; if (g > p) a1 = lgp; }
      else { di0 = inv; if (g > p) a0 = lgp; }
    }
  }
  return bad;
}

// Zeroes the tile triangle of front f in shared memory (1.0 on the padding pivots) and adds the plan's items: H and b
// from V, then the children's Schur complements / update vectors from the replica's front storage.
__device__ __forceinline__ void tile_assemble(const TileArgs& A, int f, double* F, int ntile, const double* __restrict__ V,
                                              const double* Fr, int lane, int s, int sp) {
  double2* F2 = reinterpret_cast<double2*>(F);
  // The assembly below is a chain item -> value -> add; the children's Schur complements were written by the
  // previous launches and sit in HBM.  Ask for their tile rows (rows >= the child's KT are contiguous) now, so
  // that they are on their way to L2 while the front is zeroed and the H entries are placed.
  for (int ci = A.child_ptr[f]; ci < A.child_ptr[f + 1]; ci++) {
    const int ch = A.children[ci];
    const int spc = (A.npiv[ch] + 7) & ~7, KTc = spc >> 3;
    const double* base = Fr + A.fptr[ch];
    const long lo = tile_base(KTc, 0), hi = A.fptr[ch + 1] - A.fptr[ch];
    for (long o = lo + 16 * lane; o < hi; o += 16 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
  }
  // ---- assembly: 8 items per lane in flight; the index pairs of the next batch are requested before the
  // values of the current one are used ----
  const int a0 = A.item_ptr[f], a1 = A.item_ptr[f + 1], av = a0 + A.item_nv[f];
  constexpr int IB = 8;
  int2 it[IB];
#pragma unroll
  for (int k = 0; k < IB; k++) {
    const int q = a0 + 32 * k + lane;
    it[k] = q < av ? __ldg(A.items + q) : make_int2(-1, 0);
  }
  for (int q = lane; q < (ntile << 5); q += 32) F2[q] = make_double2(0.0, 0.0);
  __syncwarp();
  if (lane < sp - s) F[tile_at(s + lane, s + lane)] = 1.0;  // padding pivots: identity
  for (int q0 = a0; q0 < av; q0 += 32 * IB) {  // H and b: destinations are distinct over the whole group
    double v[IB];
#pragma unroll
    for (int k = 0; k < IB; k++) v[k] = it[k].x >= 0 ? __ldg(V + it[k].x) : 0.0;
    int2 nx[IB];
#pragma unroll
    for (int k = 0; k < IB; k++) {
      const int q = q0 + 32 * IB + 32 * k + lane;
      nx[k] = q < av ? __ldg(A.items + q) : make_int2(-1, 0);
    }
#pragma unroll
    for (int k = 0; k < IB; k++)
      if (it[k].x >= 0) F[it[k].y] = v[k];
#pragma unroll
    for (int k = 0; k < IB; k++) it[k] = nx[k];
  }
#pragma unroll
  for (int k = 0; k < IB; k++) {
    const int q = av + 32 * k + lane;
    it[k] = q < a1 ? __ldg(A.items + q) : make_int2(-1, 0);
  }
  __syncwarp();
  for (int q0 = av; q0 < a1; q0 += 32 * IB) {  // children: the 32 items of one step never share a destination
    double v[IB];
#pragma unroll
    for (int k = 0; k < IB; k++) v[k] = it[k].x >= 0 ? Fr[it[k].x] : 0.0;
    int2 nx[IB];
#pragma unroll
    for (int k = 0; k < IB; k++) {
      const int q = q0 + 32 * IB + 32 * k + lane;
      nx[k] = q < a1 ? __ldg(A.items + q) : make_int2(-1, 0);
    }
#pragma unroll
    for (int k = 0; k < IB; k++) {
      if (it[k].x >= 0) F[it[k].y] += v[k];
      __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < IB; k++) it[k] = nx[k];
  }
}

__global__ void __launch_bounds__(128)
factor_tile_kernel(TileArgs A, int list_off, int R, int slab, const double* __restrict__ V_all, long nV, double* F_all,
                   long nF, int* status, double* x_all, int n) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // consecutive warps / CTAs = consecutive replicas of ONE front: its item list stays in L1
  const int r = blockIdx.x * (blockDim.x >> 5) + wid;
  if (r >= R) return;  // no block-wide barrier below
  const int f = A.list[list_off + blockIdx.y];
  const int s = A.npiv[f], u = A.nupd[f];
  const int sp = (s + 7) & ~7, nloc = sp + u + 1, T = (nloc + 7) >> 3, KT = sp >> 3;
  const int ntile = (T * (T + 1)) >> 1;
  double* F = smem + (size_t)wid * slab;
  double2* F2 = reinterpret_cast<double2*>(F);
  const double* V = V_all + (size_t)r * nV;
  double* Fr = F_all + (size_t)r * nF;
  tile_assemble(A, f, F, ntile, V, Fr, lane, s, sp);
  // ---- factorisation ----
  const int g = lane >> 2, t = lane & 3;
  const int sw = (g & 2) << 1;            // column-half swap of this lane's row (tileplan.h)
  const int cl = g * 8 + ((2 * t) ^ sw);  // this lane's two entries of a tile (accumulator layout)
  const int al0 = g * 8 + (t ^ sw);       // its entry of the first k-half as A / B operand: column t ...
  const int al1 = g * 8 + ((4 + t) ^ sw); // ... and of the second: column 4 + t
  bool bad = false;
  for (int K = 0; K < KT; K++) {
    double* Dk = F + tile_base(K, K);
    double di0 = 1.0, di1 = 1.0, w0, w1, nd0, nd1;
    {
      double2 dv = *reinterpret_cast<double2*>(Dk + cl);
      bad |= tile_ldlt(dv.x, dv.y, di0, di1, g, t);
      __syncwarp();
      *reinterpret_cast<double2*>(Dk + cl) = dv;  // final: unit lower L, D on the diagonal
      __syncwarp();
      // row g of inv(L11): x_j = [j == g] - sum_{k > j} x_k l_kj, j descending (x_k = 0 for k > g by itself)
      double x[8];
#pragma unroll
      for (int j = 7; j >= 0; j--) {
        double acc = (j == g) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 7; k > j; k--) acc -= x[k] * Dk[tile_in(k, j)];
        x[j] = acc;
      }
      // B operand of X = A W, W = inv(L11)^T: W[t][g] = inv(L11)[g][t], W[4+t][g] = inv(L11)[g][4+t]
      w0 = t == 0 ? x[0] : t == 1 ? x[1] : t == 2 ? x[2] : x[3];
      w1 = t == 0 ? x[4] : t == 1 ? x[5] : t == 2 ? x[6] : x[7];
      // -d of the columns t and 4 + t: turns a stored L entry back into the (negated) entry of X = L D
      nd0 = -Dk[tile_in(t, t)];
      nd1 = -Dk[tile_in(4 + t, 4 + t)];
    }
    for (int I = K + 1; I < T; I++) {
      double* P = F + tile_base(I, K);
      const double pa0 = P[al0], pa1 = P[al1];
      double x0 = 0.0, x1 = 0.0;
      dmma_acc(x0, x1, pa0, w0);  // X = A_IK W = L_IK D
      dmma_acc(x0, x1, pa1, w1);
      __syncwarp();  // all lanes hold their A-operand entries before the tile is overwritten
      *reinterpret_cast<double2*>(P + cl) = make_double2(x0 * di0, x1 * di1);  // L_IK, final
      __syncwarp();
      const double xa0 = P[al0] * nd0, xa1 = P[al1] * nd1;  // -X in A-operand layout
      const int rowI = tile_base(I, 0);
      for (int J0 = K + 1; J0 <= I; J0 += 4) {  // C_IJ -= X L_JK^T, up to four tiles in flight
        double b0[4], b1[4];
        double2 cv[4];
        const int nq = min(4, I - J0 + 1);  // warp-uniform
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (q < nq) {
            const double* Lj = F + tile_base(J0 + q, K);
            b0[q] = Lj[al0];
            b1[q] = Lj[al1];
            cv[q] = *reinterpret_cast<const double2*>(F + rowI + ((J0 + q) << 6) + cl);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (q < nq) {
            dmma_acc(cv[q].x, cv[q].y, xa0, b0[q]);
            dmma_acc(cv[q].x, cv[q].y, xa1, b1[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (q < nq) *reinterpret_cast<double2*>(F + rowI + ((J0 + q) << 6) + cl) = cv[q];
      }
      __syncwarp();
    }
  }
  __syncwarp();
  // ---- results: the tile triangle (L, Schur complement, update vector) and z ----
  double2* G2 = reinterpret_cast<double2*>(Fr + A.fptr[f]);
  for (int q = lane; q < (ntile << 5); q += 32) G2[q] = F2[q];
  double* xr = x_all + (size_t)r * n + A.piv0[f];
  const int rr = sp + u;
  for (int p = lane; p < s; p += 32) xr[p] = F[tile_at(rr, p)];
  if (bad && lane == 0) status[2 * r] = 1;
}

// The same factorisation with the WHOLE front in registers (accumulator layout: tile q = two doubles per lane),
// one instantiation per number of tile rows T <= 8.  ncu on the shared-memory-resident kernel above: L1/TEX data
// pipe 88 % busy (990 shared-memory wavefronts per front: every tile update loads and stores its C tile and
// re-reads both operands), issue slots 55 %.  Here a tile update is two MMAs on registers; shared memory is only
// the assembly buffer and the place where a freshly computed L tile changes from the accumulator layout to the
// A/B-operand layout (one store + two loads per panel tile; the same operand registers serve as the row's -X d
// and as every later row's L^T), and the front goes to HBM straight from the registers.
template <int T>
__device__ __forceinline__ void factor_tile_reg_body(const TileArgs& A, int list_off, int R, const double* __restrict__ V_all, long nV, double* F_all,
                       long nF, int* status, double* x_all, int n) {
  constexpr int NT = (T * (T + 1)) / 2;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x * (blockDim.x >> 5) + wid;
  if (r >= R) return;  // no block-wide barrier below
  const int f = A.list[list_off + blockIdx.y];  // every front of this launch has exactly T tile rows
  const int s = A.npiv[f], u = A.nupd[f];
  const int sp = (s + 7) & ~7, KT = sp >> 3, rr = sp + u;
  double* F = smem + (size_t)wid * (NT * 64);
  const double* V = V_all + (size_t)r * nV;
  double* Fr = F_all + (size_t)r * nF;
  tile_assemble(A, f, F, NT, V, Fr, lane, s, sp);
  __syncwarp();
  const int g = lane >> 2, t = lane & 3;
  const int sw = (g & 2) << 1;
  const int cl = g * 8 + ((2 * t) ^ sw), al0 = g * 8 + (t ^ sw), al1 = g * 8 + ((4 + t) ^ sw);
  double2 c[NT];
#pragma unroll
  for (int q = 0; q < NT; q++) c[q] = *reinterpret_cast<const double2*>(F + q * 64 + cl);
  __syncwarp();
  bool bad = false;
#pragma unroll
  for (int K = 0; K < T - 1; K++) {  // KT <= T - 1: the rhs row lies behind the pivots
    if (K < KT) {
      const int dkk = (K * (K + 1)) / 2 + K;
      double di0 = 1.0, di1 = 1.0;
      bad |= tile_ldlt(c[dkk].x, c[dkk].y, di0, di1, g, t);
      double* Dk = F + dkk * 64;
      *reinterpret_cast<double2*>(Dk + cl) = c[dkk];
#pragma unroll
      for (int I = K + 1; I < T; I++) *reinterpret_cast<double2*>(F + ((I * (I + 1)) / 2 + K) * 64 + cl) = c[(I * (I + 1)) / 2 + K];
      __syncwarp();
      double x[8];
#pragma unroll
      for (int j = 7; j >= 0; j--) {
        double acc = (j == g) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 7; k > j; k--) acc -= x[k] * Dk[tile_in(k, j)];
        x[j] = acc;
      }
      const double w0 = t == 0 ? x[0] : t == 1 ? x[1] : t == 2 ? x[2] : x[3];
      const double w1 = t == 0 ? x[4] : t == 1 ? x[5] : t == 2 ? x[6] : x[7];
      const double nd0 = -Dk[tile_in(t, t)], nd1 = -Dk[tile_in(4 + t, 4 + t)];
      // X = A_IK W = L_IK D.  The two MMAs of a tile depend on each other (k = 0..3, then k = 4..7 into the same
      // accumulator) and a warp issues in order: all first halves go out before any second half, so no MMA waits
      // for the one in front of it.
      {
        double pb[T];
#pragma unroll
        for (int I = K + 1; I < T; I++) {
          const double* P = F + ((I * (I + 1)) / 2 + K) * 64;
          double2& q = c[(I * (I + 1)) / 2 + K];
          pb[I] = P[al1];
          q = make_double2(0.0, 0.0);
          dmma_acc(q.x, q.y, P[al0], w0);
        }
#pragma unroll
        for (int I = K + 1; I < T; I++) {
          double2& q = c[(I * (I + 1)) / 2 + K];
          dmma_acc(q.x, q.y, pb[I], w1);
        }
#pragma unroll
        for (int I = K + 1; I < T; I++) {
          double2& q = c[(I * (I + 1)) / 2 + K];
          q = make_double2(q.x * di0, q.y * di1);  // L_IK, final
        }
      }
      __syncwarp();  // every A-operand read is done before the staged tiles are overwritten with L
#pragma unroll
      for (int I = K + 1; I < T; I++) *reinterpret_cast<double2*>(F + ((I * (I + 1)) / 2 + K) * 64 + cl) = c[(I * (I + 1)) / 2 + K];
      __syncwarp();
      double la0[T], la1[T];  // L_IK in operand layout: B operand of every row >= I, and (times -d) A operand of row I
#pragma unroll
      for (int I = K + 1; I < T; I++) {
        const double* P = F + ((I * (I + 1)) / 2 + K) * 64;
        la0[I] = P[al0];
        la1[I] = P[al1];
      }
#pragma unroll
      for (int I = K + 1; I < T; I++) {  // first k-half of every tile update, then the second (see above)
        const double xa0 = la0[I] * nd0;
#pragma unroll
        for (int J = K + 1; J <= I; J++) dmma_acc(c[(I * (I + 1)) / 2 + J].x, c[(I * (I + 1)) / 2 + J].y, xa0, la0[J]);
      }
#pragma unroll
      for (int I = K + 1; I < T; I++) {
        const double xa1 = la1[I] * nd1;
#pragma unroll
        for (int J = K + 1; J <= I; J++) dmma_acc(c[(I * (I + 1)) / 2 + J].x, c[(I * (I + 1)) / 2 + J].y, xa1, la1[J]);
      }
    }
  }
  // ---- results straight from the registers: the tile triangle (16-byte stores, 512 bytes per tile) and z ----
  double* Fg = Fr + A.fptr[f];
#pragma unroll
  for (int q = 0; q < NT; q++) *reinterpret_cast<double2*>(Fg + q * 64 + cl) = c[q];
  double* xr = x_all + (size_t)r * n + A.piv0[f];
  if (g == (rr & 7)) {  // the rhs row lives in tile row T - 1
#pragma unroll
    for (int K = 0; K < T - 1; K++) {
      if (K < KT) {
        const double2 z = c[((T - 1) * T) / 2 + K];
        if (8 * K + 2 * t < s) xr[8 * K + 2 * t] = z.x;
        if (8 * K + 2 * t + 1 < s) xr[8 * K + 2 * t + 1] = z.y;
      }
    }
  }
  if (bad && lane == 0) status[2 * r] = 1;
}

template <int T>
__global__ void __launch_bounds__(128)
factor_tile_reg_kernel(TileArgs A, int list_off, int R, const double* __restrict__ V_all, long nV, double* F_all,
                       long nF, int* status, double* x_all, int n) {
  factor_tile_reg_body<T>(A, list_off, R, V_all, nV, F_all, nF, status, x_all, n);
}
// Register caps for more resident warps: the kernel is latency-bound at the 12-16 warps per SM the compiler's own
// register count allows, and spilling a few dozen registers costs less than the extra warps bring (measured, C3:
// 415k -> 425k replica GN it/s with the level-1 caps).  One warp per CTA.  SLAM_B200_TILE_TIGHT = 0 (none) ... 3.
constexpr int tile_reg_min_blocks(int T, int lvl) {
  return lvl == 1 ? (T <= 5 ? 20 : T == 6 ? 18 : T == 7 ? 14 : 9)
       : lvl == 2 ? (T <= 5 ? 24 : T == 6 ? 20 : T == 7 ? 16 : 11)
                  : (T <= 5 ? 28 : T == 6 ? 24 : T == 7 ? 18 : 12);
}
template <int T, int LVL>
__global__ void __launch_bounds__(32, tile_reg_min_blocks(T, LVL))
factor_tile_reg_tight_kernel(TileArgs A, int list_off, int R, const double* __restrict__ V_all, long nV, double* F_all,
                             long nF, int* status, double* x_all, int n) {
  factor_tile_reg_body<T>(A, list_off, R, V_all, nV, F_all, nF, status, x_all, n);
}

// Backward sweep (L^T x = z, root -> leaves) from the stored L tiles: per pivot tile column all tiles below
// and the diagonal tile are requested at once (16-byte loads, unit stride), multiplied with the already-known
// x of their rows and reduced over the 8 rows of the accumulator layout by shuffles; the 8 x 8 unit triangle
// goes through shared memory and is solved redundantly by every lane.
constexpr int TILE_MAX_T = TILE_MAX_ROWS / 8;

constexpr int BACK_SLAB = TILE_MAX_ROWS + 8 + 64;  // x of the front's rows + the current unit triangle, per warp

// TMAXT: the most tile rows any front of the launch has (8 for the trackdrive topologies: 28 instead of 44 registers of
// tile buffers, 24 instead of 20 resident warps per SM)
template <int TMAXT>
__global__ void __launch_bounds__(128, TMAXT <= 8 ? 6 : 5)
backward_tile_kernel(TileArgs A, int list_off, int R, const double* __restrict__ F_all, long nF, double* x_all, int n,
                     int early) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.x * (blockDim.x >> 5) + wid;
  if (r >= R) return;
  const int f = A.list[list_off + blockIdx.y];
  double* tri = smem + (size_t)wid * BACK_SLAB;  // 64 doubles, 16-byte aligned (BACK_SLAB is even)
  double* xs = tri + 64;
  const int s = A.npiv[f], u = A.nupd[f], p0 = A.piv0[f];
  const int sp = (s + 7) & ~7, nloc = sp + u + 1, T = (nloc + 7) >> 3, KT = sp >> 3;
  double* x = x_all + (size_t)r * n;
  const int* rows = A.upd_rows + A.rows_ptr[f];
  const double* Fg = F_all + (size_t)r * nF + A.fptr[f];
  const int g = lane >> 2, t = lane & 3;
  const int cl = g * 8 + ((2 * t) ^ ((g & 2) << 1));  // this lane's two entries of a tile (accumulator layout, tileplan.h)
  // Programmatic dependent launch (pdl_wait_then_release): the levels of the backward sweep are one launch each on one
  // stream; with early set (every launch but the first, whose predecessor is the root's factor launch) the index chain
  // and the first tiles of L -- written by the factor launches, long complete -- are fetched under the previous level.
  if (!early) pdl_wait_then_release();
  // the last pivot column's tiles do not depend on x: request them before the gather of x
  double2 lv[TMAXT - 1], dg;
  {
    const int K = KT - 1;
    dg = *reinterpret_cast<const double2*>(Fg + tile_base(K, K) + cl);
#pragma unroll
    for (int d = 0; d < TMAXT - 1; d++)
      if (K + 1 + d < T) lv[d] = *reinterpret_cast<const double2*>(Fg + tile_base(K + 1 + d, K) + cl);
  }
  if (early) pdl_wait_then_release();  // the parents' x comes from the previous launch
  for (int i = lane; i < (T << 3); i += 32) {
    double v = 0.0;  // padding pivots, the rhs row and the rows behind it contribute nothing
    if (i < s) v = __ldcg(x + p0 + i);
    else if (i >= sp && i < sp + u) v = __ldcg(x + rows[i - sp]);
    xs[i] = v;
  }
  __syncwarp();
  for (int K = KT - 1; K >= 0; K--) {
    *reinterpret_cast<double2*>(tri + cl) = dg;
    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
    for (int d = 0; d < TMAXT - 1; d++) {
      if (K + 1 + d < T) {
        const double xi = xs[8 * (K + 1 + d) + g];
        acc0 += lv[d].x * xi;
        acc1 += lv[d].y * xi;
      }
    }
    if (K > 0) {  // the next column's tiles while this one is reduced and solved
      dg = *reinterpret_cast<const double2*>(Fg + tile_base(K - 1, K - 1) + cl);
#pragma unroll
      for (int d = 0; d < TMAXT - 1; d++)
        if (K + d < T) lv[d] = *reinterpret_cast<const double2*>(Fg + tile_base(K + d, K - 1) + cl);
    }
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
    }
    if (g == 0) {
      xs[8 * K + 2 * t] -= acc0;
      xs[8 * K + 2 * t + 1] -= acc1;
    }
    __syncwarp();
    double w[8];
#pragma unroll
    for (int c = 0; c < 8; c++) w[c] = xs[8 * K + c];
#pragma unroll
    for (int i = 7; i >= 1; i--)
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (j < i) w[j] -= tri[tile_in(i, j)] * w[i];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; c++)
      if (lane == c) xs[8 * K + c] = w[c];
    __syncwarp();
  }
  for (int i = lane; i < s; i += 32) x[p0 + i] = xs[i];
}

#include "factor3.cuh"

// L panel (fs x s, leading dimension fs) from global into shared memory with leading dimension ld:
// one warp per column, lanes over rows, up to five independent loads in flight per lane
__device__ __forceinline__ void stage_panel(double* Ls, const double* __restrict__ Lg, int fs, int s, int ld,
                                            int warp, int nw, int lane) {
  for (int j = warp; j < s; j += nw) {
    const double* src = Lg + (size_t)j * fs;
    double* dst = Ls + (size_t)j * ld;
    for (int i = j + lane; i < fs; i += 160) {  // rows above the diagonal are never read
      const int i1 = i + 32, i2 = i + 64, i3 = i + 96, i4 = i + 128;
      const double v0 = src[i];
      const double v1 = i1 < fs ? src[i1] : 0.0;
      const double v2 = i2 < fs ? src[i2] : 0.0;
      const double v3 = i3 < fs ? src[i3] : 0.0;
      const double v4 = i4 < fs ? src[i4] : 0.0;
      dst[i] = v0;
      if (i1 < fs) dst[i1] = v1;
      if (i2 < fs) dst[i2] = v2;
      if (i3 < fs) dst[i3] = v3;
      if (i4 < fs) dst[i4] = v4;
    }
  }
}

// ---- forward solve: L y = b, then z = D^-1 y, one front per CTA, leaves -> root ------------------
// Blocked like the factorisation: per panel of NB pivots every thread solves the NB x NB unit
// triangle redundantly in registers, then one thread per remaining row subtracts the panel's
// contribution.  One barrier per NB pivots.  The L panel is staged in shared memory with an odd
// leading dimension (conflict-free row and column access).
template <bool SMEM>
__global__ void __launch_bounds__(SOLVE_THREADS)
forward_kernel(SymArgs S, int list_off, const double* __restrict__ V_all, long nV,
               const double* __restrict__ Lv_all, long nL, double* uvec_all, long nUvec, double* x_all, int n) {
  extern __shared__ double smem[];
  const int g = S.launch_list[list_off + blockIdx.x];
  const int r = blockIdx.y;
  const int s = S.npiv[g], u = S.nupd[g], fs = s + u, p0 = S.piv0[g];
  const double* V = V_all + (size_t)r * nV;
  const double* Lg = Lv_all + (size_t)r * nL + S.lptr[g];
  double* uvec = uvec_all + (size_t)r * nUvec;
  double* x = x_all + (size_t)r * n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int ld = SMEM ? (fs | 1) : fs;
  double* w = smem;             // fs
  double* ys = smem + fs;       // s (solution of the unit-triangular part)
  double* Ls = smem + fs + s;   // ld * s when SMEM
  // w = rhs of the pivots + the children's update vectors, gathered per destination row in child
  // order (fixed summation order, no atomics, no per-child barrier)
  const int* gp = S.gather_ptr + S.frow_ptr[g];
  for (int i = tid; i < fs; i += nt) {
    double acc = i < s ? V[S.solver2v[p0 + i]] : 0.0;
    for (int q = gp[i]; q < gp[i + 1]; q++) acc += uvec[S.gather_src[q]];
    w[i] = acc;
  }
  if (SMEM) stage_panel(Ls, Lg, fs, s, ld, warp, nw, lane);
  __syncthreads();
  const double* Lp = SMEM ? Ls : Lg;
  for (int k0 = 0; k0 < s; k0 += NB) {
    const int nb = min(NB, s - k0);
    double y[NB];
#pragma unroll
    for (int p = 0; p < NB; p++) y[p] = (p < nb) ? w[k0 + p] : 0.0;
#pragma unroll
    for (int p = 0; p < NB; p++)
#pragma unroll
      for (int q = p + 1; q < NB; q++)
        if (q < nb) y[q] -= Lp[(size_t)(k0 + p) * ld + k0 + q] * y[p];
    for (int i = k0 + nb + tid; i < fs; i += nt) {
      double acc = w[i];
#pragma unroll
      for (int p = 0; p < NB; p++)
        if (p < nb) acc -= Lp[(size_t)(k0 + p) * ld + i] * y[p];
      w[i] = acc;
    }
    if (tid < nb) {
#pragma unroll
      for (int p = 0; p < NB; p++)
        if (p == tid) ys[k0 + p] = y[p];
    }
    __syncthreads();
  }
  for (int i = tid; i < fs; i += nt) {
    if (i < s) x[p0 + i] = ys[i] / Lp[(size_t)i * ld + i];
    else uvec[S.rows_ptr[g] + i - s] = w[i];
  }
}

// ---- backward solve: L^T x = z, root -> leaves ----------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(SOLVE_THREADS)
backward_kernel(SymArgs S, int list_off, const double* __restrict__ Lv_all, long nL, double* x_all, int n, int early) {
  extern __shared__ double smem[];
  const int g = S.launch_list[list_off + blockIdx.x];
  const int r = blockIdx.y;
  const int s = S.npiv[g], u = S.nupd[g], fs = s + u, p0 = S.piv0[g];
  const double* Lg = Lv_all + (size_t)r * nL + S.lptr[g];
  double* x = x_all + (size_t)r * n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  const int ld = SMEM ? (fs | 1) : fs;
  double* xs = smem;            // fs
  double* xo = smem + fs;       // s
  double* Ls = smem + fs + s;
  const int* rows = S.upd_rows + S.rows_ptr[g];
  // early: this front's L panel was written at least two launches back and is staged while the previous backward
  // launch still runs (the first backward launch follows the root's factor launch directly: early = 0)
  const bool tlc = S.tl && r == 0 && tid == 0;
  if (tlc) S.tl[TLS * g + 6] = global_ns();
  if (!early) pdl_wait_then_release();
  if (SMEM) stage_panel(Ls, Lg, fs, s, ld, warp, nw, lane);
  const int xi0 = tid < fs ? (tid < s ? p0 + tid : rows[tid - s]) : 0;  // static: fetched before the wait
  if (early) pdl_wait_then_release();
  if (tlc) S.tl[TLS * g + 7] = global_ns();
  if (tid < fs) xs[tid] = __ldcg(x + xi0);
  for (int i = tid + nt; i < fs; i += nt) xs[i] = __ldcg(i < s ? x + p0 + i : x + rows[i - s]);
  __syncthreads();
  const double* Lp = SMEM ? Ls : Lg;
  // contribution of the already-solved ancestor rows: xs[k] -= sum_{i>=s} L[i,k] xs[i]
  for (int k = warp; k < s; k += nw) {
    const double* col = Lp + (size_t)k * ld;
    double t = 0.0;
    for (int i = s + lane; i < fs; i += 32) t += col[i] * xs[i];
    for (int o = 16; o; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) xs[k] -= t;
  }
  __syncthreads();
  // unit upper-triangular solve with L11^T, panels from the last pivot down: warp 0 solves the
  // NB x NB triangle (a dependent chain), everybody then subtracts the panel from the earlier rows
  const int npan = (s + NB - 1) / NB;
  if (SMEM && s <= 64) {
    // At most 64 pivots: ONE warp holds the whole pivot part of x (two entries per lane) and walks the columns from the
    // last pivot down -- x_k is final, every earlier entry takes its L[k][i] x_k -- by shuffles, without a block barrier
    // (the panel loop below: two barriers and a 28-FMA triangle per 8 pivots; 7 us -> see RESULTS for a 46-pivot front).
    if (warp == 0) {
      const int i0 = lane, i1 = lane + 32;
      double x0 = i0 < s ? xs[i0] : 0.0, x1 = i1 < s ? xs[i1] : 0.0;
      const double* c0p = Lp + (size_t)i0 * ld;  // column i0 of L: L[k][i0] = c0p[k]
      const double* c1p = Lp + (size_t)i1 * ld;
      for (int k = s - 1; k > 0; k--) {
        const double l0 = i0 < k ? c0p[k] : 0.0, l1 = i1 < k ? c1p[k] : 0.0;
        const double xk = __shfl_sync(0xffffffffu, k < 32 ? x0 : x1, k & 31);
        x0 -= l0 * xk;
        x1 -= l1 * xk;
      }
      if (i0 < s) xo[i0] = x0;
      if (i1 < s) xo[i1] = x1;
    }
    __syncthreads();
  } else
  for (int pan = npan - 1; pan >= 0; pan--) {
    const int k0 = pan * NB;
    const int nb = min(NB, s - k0);
    if (warp == 0) {
      double xp[NB];
#pragma unroll
      for (int p = 0; p < NB; p++) xp[p] = (p < nb) ? xs[k0 + p] : 0.0;
#pragma unroll
      for (int p = NB - 1; p >= 0; p--)
#pragma unroll
        for (int q = 0; q < p; q++)
          if (p < nb) xp[q] -= Lp[(size_t)(k0 + q) * ld + k0 + p] * xp[p];
      if (lane == 0) {
#pragma unroll
        for (int p = 0; p < NB; p++)
          if (p < nb) xo[k0 + p] = xp[p];
      }
    }
    __syncthreads();
    for (int k = tid; k < k0; k += nt) {
      const double* col = Lp + (size_t)k * ld + k0;
      double acc = xs[k];
#pragma unroll
      for (int p = 0; p < NB; p++)
        if (p < nb) acc -= col[p] * xo[k0 + p];
      xs[k] = acc;
    }
    __syncthreads();
  }
  for (int i = tid; i < s; i += nt) x[p0 + i] = xo[i];
  if (tlc) S.tl[TLS * g + 8] = global_ns();
}

// Small fronts (<= 64 rows): one warp per front when there are enough of them to fill the machine
// (batched replicas, very large graphs), otherwise one 128-thread CTA per front (a single mid-size
// graph has only ~1,000 leaves: more threads per front shorten the level's latency).
bool warp_kernels(const slam_b200_ctx* c, const DeviceSystem& D, const LevelLaunch& LL) {
  return (long)LL.n_tiny * D.R >= 16L * c->num_sms;
}

size_t factor_extra_smem(int max_fs) {  // srel (ints, even count) + scaled panel
  return (size_t)((max_fs + 1) & ~1) * sizeof(int) + ((size_t)(NB + 2) * max_fs + 8 * NB + 2 * (NB * NB + 2 * NB)) * sizeof(double);
}

// Launch with programmatic stream serialisation (see pdl_wait_then_release): the kernel may become resident as soon as
// every CTA of the launch before it on the stream has released its dependents.  Captured into the iteration's CUDA
// graph the attribute becomes a programmatic dependency edge.
bool pdl_enabled() {
  static const bool on = getenv("SLAM_B200_NO_PDL") == nullptr;
  return on;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_front(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

SymArgs sym_args(const DeviceSystem& D) {
  SymArgs a;
  a.piv0 = D.ds.piv0.p; a.npiv = D.ds.npiv.p; a.nupd = D.ds.nupd.p; a.rows_ptr = D.ds.rows_ptr.p;
  a.upd_rows = D.ds.upd_rows.p; a.rel = D.ds.rel.p; a.child_ptr = D.ds.child_ptr.p;
  a.children = D.ds.children.p; a.asm_ptr = D.ds.asm_ptr.p; a.solver2v = D.ds.solver2v.p;
  a.lptr = D.ds.lptr.p; a.uptr = D.ds.uptr.p; a.fbig = D.ds.fbig.p;
  a.asm_entries = D.ds.asm_entries.p; a.launch_list = D.ds.launch_list.p;
  a.frow_ptr = D.ds.frow_ptr.p; a.gather_ptr = D.ds.gather_ptr.p; a.gather_src = D.ds.gather_src.p;
  a.dbg = D.dbg_clocks.p;
  a.tl = D.timeline.p;
  static const int dbg_front = getenv("SLAM_B200_DBG_FRONT") ? atoi(getenv("SLAM_B200_DBG_FRONT")) : -1;
  a.dbg_front = dbg_front;
  return a;
}

}  // namespace

int graph_enqueue_update(slam_b200_ctx* c);  // graph.cu

// Opt-in shared-memory limits of the front kernels.  Function attributes belong to the device the
// call is made on, so the flag lives in the context (a process may hold contexts on several devices)
// and not in a process-wide global.  Always reached outside stream capture first
// (graph_enqueue_iteration calls it before cudaStreamBeginCapture).
static int solver_init_attrs(slam_b200_ctx* c) {
  if (!c->solver_attrs_set) {
    int lim = c->max_smem_optin;
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<5, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<6, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<6, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<6, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<7, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<7, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    SLAM_CUDA_TRY(c, cudaFuncSetAttribute(factor_tile_reg_tight_kernel<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    c->solver_attrs_set = true;
  }
  return 0;
}

int graph_enqueue_solve(slam_b200_ctx* c) {
  DeviceSystem& D = *c->sys;
  if (int rc = solver_init_attrs(c)) return rc;
  SymArgs S = sym_args(D);
  const size_t smem_limit = (size_t)std::max(0, c->max_smem_optin - 1024);
  const int nlv = (int)D.levels.size();
  auto mark = [&]() {
    if (!D.profile) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, c->stream);
    D.prof_events.push_back(e);
  };
  if (D.tile_path) {
    TileArgs TA;
    TA.npiv = D.ds.npiv.p; TA.nupd = D.ds.nupd.p; TA.piv0 = D.ds.piv0.p; TA.rows_ptr = D.ds.rows_ptr.p;
    TA.upd_rows = D.ds.upd_rows.p; TA.list = D.tile_list.p; TA.item_ptr = D.tile_item_ptr.p;
    TA.item_nv = D.tile_item_nv.p; TA.fptr = D.tile_fptr.p; TA.items = D.tile_items.p;
    TA.child_ptr = D.ds.child_ptr.p; TA.children = D.ds.children.p;
    // one launch per (level, class of tile rows) sizes the shared memory for its class; a level with less
    // than two waves of CTAs is launch- and latency-bound instead, so its classes go out as ONE launch.
    // grid = (replica, front): consecutive CTAs are consecutive replicas of one front.
    const long small_level = 2L * c->num_sms * 16;
    const int nTL = (int)D.tile_launches.size();
    // warps (= replicas) per CTA: measurement switches SLAM_B200_TILE_WPC_F / _B (1, 2 or 4)
    static const int wpc_f = [] { const char* e = getenv("SLAM_B200_TILE_WPC_F"); int v = e ? atoi(e) : 1; return (v == 1 || v == 2 || v == 4) ? v : 1; }();
    static const int wpc_b = [] { const char* e = getenv("SLAM_B200_TILE_WPC_B"); int v = e ? atoi(e) : 4; return (v == 1 || v == 2 || v == 4) ? v : 4; }();
    static const bool reg_path = getenv("SLAM_B200_TILE_NO_REG") == nullptr;
    // level of the register caps per number of tile rows (tile_reg_min_blocks): one digit for all T, or one digit each
    // for T = 5, 6, 7, 8 (e.g. SLAM_B200_TILE_TIGHT=2110)
    static const std::array<int, 9> tight_of = [] {
      std::array<int, 9> a;
      a.fill(1);
      if (const char* e = getenv("SLAM_B200_TILE_TIGHT")) {
        const size_t len = strlen(e);
        for (int T = 0; T <= 8; T++) {
          const size_t k = len == 1 ? 0 : (T >= 5 ? (size_t)(T - 5) : 0);
          const char ch = k < len ? e[k] : '1';
          a[T] = (ch >= '0' && ch <= '3') ? ch - '0' : 1;
        }
      }
      return a;
    }();
    // The size classes of one level touch disjoint fronts, so they may run side by side: class i > 0 goes to a side
    // stream between a fork and a join event (parallel branches once the iteration is captured into a CUDA graph).
    // Levels near the root hold one to three fronts per class -- less than a wave of warps each -- and would otherwise
    // pay one front latency per class.  SLAM_B200_TILE_NO_FORK=1 keeps everything on the context's stream.
    static const bool fork_classes = getenv("SLAM_B200_TILE_NO_FORK") == nullptr;
    size_t ev_used = 0;
    auto next_event = [&]() -> cudaEvent_t {
      if (ev_used == c->fork_events.size()) {
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        c->fork_events.push_back(e);
      }
      return c->fork_events[ev_used++];
    };
    auto side_stream = [&](int i) -> cudaStream_t {
      cudaStream_t& a = c->aux_stream[i % slam_b200_ctx::kAuxStreams];
      if (!a) cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking);
      return a;
    };
    auto launch_factor = [&](cudaStream_t st, int list_off, int count, int Tmax, bool uniform_T) {
      const int slab = ((Tmax * (Tmax + 1)) / 2) * 64 + TILE_SCRATCH;
      if (reg_path && uniform_T && Tmax <= 8) {  // whole front in registers, one instantiation per T
        const int tight = tight_of[Tmax];
        for (int o = 0; o < count; o += 65535) {
          dim3 grid((D.R + wpc_f - 1) / wpc_f, std::min(65535, count - o));
          const size_t sm = (size_t)wpc_f * slab * sizeof(double);
#define REG_ARGS (TA, list_off + o, D.R, D.V.p, D.nV, D.Lv.p, D.nL, D.status.p, D.x.p, D.n)
#define REG_LAUNCH(TT) case TT: \
            if (wpc_f != 1 || tight == 0) factor_tile_reg_kernel<TT><<<grid, 32 * wpc_f, sm, st>>> REG_ARGS; \
            else if (tight == 1) factor_tile_reg_tight_kernel<TT, 1><<<grid, 32, sm, st>>> REG_ARGS; \
            else if (tight == 2) factor_tile_reg_tight_kernel<TT, 2><<<grid, 32, sm, st>>> REG_ARGS; \
            else factor_tile_reg_tight_kernel<TT, 3><<<grid, 32, sm, st>>> REG_ARGS; \
            break
          switch (Tmax) { REG_LAUNCH(1); REG_LAUNCH(2); REG_LAUNCH(3); REG_LAUNCH(4); REG_LAUNCH(5); REG_LAUNCH(6); REG_LAUNCH(7); REG_LAUNCH(8); default: break; }
#undef REG_LAUNCH
#undef REG_ARGS
          c->launches++;
        }
        return;
      }
      for (int o = 0; o < count; o += 65535) {
        dim3 grid((D.R + wpc_f - 1) / wpc_f, std::min(65535, count - o));
        factor_tile_kernel<<<grid, 32 * wpc_f, (size_t)wpc_f * slab * sizeof(double), st>>>(
            TA, list_off + o, D.R, slab, D.V.p, D.nV, D.Lv.p, D.nL, D.status.p, D.x.p, D.n);
        c->launches++;
      }
    };
    for (int k = 0; k < nTL;) {
      int k1 = k + 1;
      while (k1 < nTL && D.tile_launches[k1].level == D.tile_launches[k].level) k1++;
      long ctas = 0;
      for (int q = k; q < k1; q++) ctas += (long)D.tile_launches[q].count * D.R;
      // classes of <= 8 tile rows hold exactly one T (graph.cu) and run the register-resident kernel; the others
      // share the shared-memory-resident kernel and may be merged
      bool all_big = true;
      for (int q = k; q < k1; q++) all_big = all_big && (D.tile_launches[q].T > 8 || !reg_path);
      const int step = (all_big && ctas <= small_level) ? k1 - k : 1;
      const bool fork = fork_classes && (k1 - k + step - 1) / step > 1;
      cudaEvent_t ev_fork = nullptr;
      if (fork) {
        ev_fork = next_event();
        SLAM_CUDA_TRY(c, cudaEventRecord(ev_fork, c->stream));
      }
      int branch = 0;
      std::vector<cudaStream_t> used;
      for (int q = k; q < k1; q += step, branch++) {
        int count = 0, Tmax = 0;
        for (int q2 = q; q2 < q + step; q2++) { count += D.tile_launches[q2].count; Tmax = std::max(Tmax, D.tile_launches[q2].T); }
        cudaStream_t st = c->stream;
        if (fork && branch > 0) {
          st = side_stream(branch - 1);
          if (branch - 1 < slam_b200_ctx::kAuxStreams) {  // first use of this side stream in this level
            SLAM_CUDA_TRY(c, cudaStreamWaitEvent(st, ev_fork, 0));
            used.push_back(st);
          }
        }
        launch_factor(st, D.tile_launches[q].list_off, count, Tmax, step == 1);
      }
      for (cudaStream_t st : used) {  // join
        cudaEvent_t ej = next_event();
        SLAM_CUDA_TRY(c, cudaEventRecord(ej, st));
        SLAM_CUDA_TRY(c, cudaStreamWaitEvent(c->stream, ej, 0));
      }
      k = k1;
    }
    mark();  // factored
    mark();  // forward (fused)
    int tile_early_l = 0;  // programmatic dependent launch along the backward sweep (backward_tile_kernel)
    for (int k1 = nTL; k1 > 0;) {  // root -> leaves, one launch per level (no dynamic shared memory: no classes)
      int k = k1 - 1;
      while (k > 0 && D.tile_launches[k - 1].level == D.tile_launches[k1 - 1].level) k--;
      int count = 0;
      for (int q = k; q < k1; q++) count += D.tile_launches[q].count;
      for (int o = 0; o < count; o += 65535) {
        dim3 grid((D.R + wpc_b - 1) / wpc_b, std::min(65535, count - o));
        if (D.tile.max_T <= 8)
          SLAM_CUDA_TRY(c, launch_front(backward_tile_kernel<8>, grid, dim3(32 * wpc_b), (size_t)wpc_b * BACK_SLAB * sizeof(double), c->stream,
              TA, D.tile_launches[k].list_off + o, D.R, D.Lv.p, D.nL, D.x.p, D.n, tile_early_l));
        else
          SLAM_CUDA_TRY(c, launch_front(backward_tile_kernel<TILE_MAX_T>, grid, dim3(32 * wpc_b), (size_t)wpc_b * BACK_SLAB * sizeof(double), c->stream,
              TA, D.tile_launches[k].list_off + o, D.R, D.Lv.p, D.nL, D.x.p, D.n, tile_early_l));
        tile_early_l = 1;
        c->launches++;
      }
      k1 = k;
    }
    SLAM_CUDA_TRY(c, cudaGetLastError());
    mark();  // backward done
    int rc = graph_enqueue_update(c);
    mark();  // updated
    return rc;
  }
  // SLAM_B200_FACTOR_VARIANT=1 selects the first-generation CTA-per-front kernel (A/B measurements)
  static const bool gen2 = !(getenv("SLAM_B200_FACTOR_VARIANT") && atoi(getenv("SLAM_B200_FACTOR_VARIANT")) == 1);
  // SLAM_B200_FACTOR_VARIANT=3: large fronts by the register-resident factor3_kernel (measured: no faster, factor3.cuh)
  static const bool gen3 = getenv("SLAM_B200_FACTOR_VARIANT") && atoi(getenv("SLAM_B200_FACTOR_VARIANT")) == 3;
  static const int la_idle = getenv("SLAM_B200_LA_IDLE") ? atoi(getenv("SLAM_B200_LA_IDLE")) : 0;
  static const bool merge_ok = getenv("SLAM_B200_NO_LEVEL_MERGE") == nullptr;
  static const int mma_flag = ((getenv("SLAM_B200_UPDATE_MMA") ? atoi(getenv("SLAM_B200_UPDATE_MMA")) : 1) ? 4 : 0) |
                              (getenv("SLAM_B200_EA_SINGLE") ? 16 : 0);  // 16: single-buffer extend-add (A/B)
  // programmatic dependent launch: 0 until a launch of this enqueue has gone out whose wait proves that the assembly
  // kernels (V) / the factor kernels (L) are complete -- see pdl_wait_then_release
  int early_v = 0, early_l = 0;
  // the last level holds roots only (fronts without update rows): their backward solve is done by the factor kernel
  // while the front is still in shared memory, and the level's backward launch is dropped (factor2_kernel, flags & 8)
  bool fuse_roots = false;
  if (gen2 && !gen3 && D.R == 1 && nlv > 0 && !getenv("SLAM_B200_SEPARATE_FORWARD") && !getenv("SLAM_B200_NO_ROOT_FUSE")) {
    const LevelLaunch& LL = D.levels[nlv - 1];
    fuse_roots = LL.n_big == 0 && !(LL.n_tiny && warp_kernels(c, D, LL));
    for (int q = 0; fuse_roots && q < LL.n_tiny + LL.n_small; q++)
      if (D.sym.nupd[D.launch_list_host[LL.list_off + q]] != 0) fuse_roots = false;
  }
  // A level that holds both fronts of <= 64 rows and larger ones would go out as two dependent launches (128- and
  // 512-thread CTAs), the second waiting for the first although they are independent (front timeline of the 10-lap
  // graph: +12 us on each of its two mixed levels).  When the whole level fits one wave of 512-thread CTAs it is ONE
  // launch of those; the launch list holds the small fronts first, by ascending size, and is walked backwards.
  auto merge_classes = [&](const LevelLaunch& LL) {
    return gen2 && !gen3 && merge_ok && D.R == 1 && LL.n_tiny > 0 && LL.n_small > 0 && !warp_kernels(c, D, LL) &&
           LL.n_tiny + LL.n_small <= 2 * c->num_sms;
  };
  for (int lv = 0; lv < nlv; lv++) {
    const LevelLaunch& LL = D.levels[lv];
    const int root_flag = (fuse_roots && lv == nlv - 1) ? 8 : 0;
    if (LL.n_tiny && warp_kernels(c, D, LL)) {
      int off = LL.list_off;
      for (int cls = 0; cls < 4; cls++) {
        const int cn = LL.tiny_cls_n[cls];
        if (!cn) continue;
        dim3 grid((cn + TINY_WARPS - 1) / TINY_WARPS, D.R);
        const int slab = (LL.tiny_cls_fs[cls] * (LL.tiny_cls_fs[cls] + 1)) / 2;
        factor_tiny_kernel<<<grid, TINY_WARPS * 32, (size_t)TINY_WARPS * slab * sizeof(double), c->stream>>>(
            S, off, cn, slab, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.status.p, D.uvec.p, D.nUvec, D.x.p, D.n);
        c->launches++;
        off += cn;
      }
    } else if (LL.n_tiny && merge_classes(LL)) {
      // handled by the 512-thread launch below
    } else if (LL.n_tiny) {
      dim3 grid(LL.n_tiny, D.R);
      if (gen2) {
        SLAM_CUDA_TRY(c, launch_front(factor2_kernel<true>, grid, dim3(128), LL.smem_tiny + factor_extra_smem(64), c->stream,
            S, LL.list_off, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.Fbig.p, D.nFbig, D.status.p, D.uvec.p, D.nUvec,
            D.x.p, D.n, la_idle, early_v | 2 | mma_flag | root_flag));
        early_v = 1;
      } else
        factor_kernel<true><<<grid, 128, LL.smem_tiny + factor_extra_smem(64), c->stream>>>(
            S, LL.list_off, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.Fbig.p, D.nFbig, D.status.p, D.uvec.p, D.nUvec,
            D.x.p, D.n);
      c->launches++;
    }
    if (LL.n_small && merge_classes(LL)) {
      dim3 grid(LL.n_tiny + LL.n_small, D.R);
      // 512 threads x 128 registers are a whole register file: one CTA per SM.  A mixed level with more fronts than SMs
      // runs 256-thread CTAs (two per SM) instead of two waves of 512-thread ones.
      static const int wide_threads = getenv("SLAM_B200_MERGE_THREADS") ? atoi(getenv("SLAM_B200_MERGE_THREADS")) : 256;
      const int threads = (LL.n_tiny + LL.n_small > c->num_sms) ? wide_threads : FACTOR_THREADS;
      SLAM_CUDA_TRY(c, launch_front(factor2_kernel<true>, grid, dim3(threads), LL.smem_factor + factor_extra_smem(LL.max_fs), c->stream,
          S, LL.list_off, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.Fbig.p, D.nFbig, D.status.p, D.uvec.p,
          D.nUvec, D.x.p, D.n, la_idle, early_v | 2 | mma_flag | root_flag));
      early_v = 1;
      c->launches++;
    } else if (LL.n_small) {
      dim3 grid(LL.n_small, D.R);
      // register-resident 16-warp kernel when every front of the class fits its 20 x 20 tile grid
      int max_nloc = 0;
      for (int q = LL.n_tiny; q < LL.n_tiny + LL.n_small; q++) {
        const int f = D.launch_list_host[LL.list_off + q];
        max_nloc = std::max(max_nloc, f3_local_rows(D.sym.npiv[f], D.sym.nupd[f]));
      }
      if (gen3 && max_nloc <= 8 * F3_MAX_T && f3_smem_bytes(max_nloc) <= smem_limit) {
        factor3_kernel<<<grid, F3_THREADS, f3_smem_bytes(max_nloc), c->stream>>>(
            S, LL.list_off + LL.n_tiny, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.status.p, D.uvec.p, D.nUvec, D.x.p, D.n);
      } else if (gen2) {
        SLAM_CUDA_TRY(c, launch_front(factor2_kernel<true>, grid, dim3(FACTOR_THREADS), LL.smem_factor + factor_extra_smem(LL.max_fs), c->stream,
            S, LL.list_off + LL.n_tiny, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.Fbig.p, D.nFbig, D.status.p, D.uvec.p,
            D.nUvec, D.x.p, D.n, la_idle, early_v | mma_flag | root_flag));
        early_v = 1;
      } else
        factor_kernel<true><<<grid, FACTOR_THREADS, LL.smem_factor + factor_extra_smem(LL.max_fs), c->stream>>>(
            S, LL.list_off + LL.n_tiny, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.Fbig.p, D.nFbig, D.status.p, D.uvec.p,
            D.nUvec, D.x.p, D.n);
      c->launches++;
    }
    if (LL.n_big) {
      dim3 grid(LL.n_big, D.R);
      if (gen2) {
        SLAM_CUDA_TRY(c, launch_front(factor2_kernel<false>, grid, dim3(FACTOR_THREADS), factor_extra_smem(LL.max_fs), c->stream,
            S, LL.list_off + LL.n_tiny + LL.n_small, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.Fbig.p, D.nFbig,
            D.status.p, D.uvec.p, D.nUvec, D.x.p, D.n, la_idle, early_v | mma_flag | root_flag));
        early_v = 1;
      } else
        factor_kernel<false><<<grid, FACTOR_THREADS, factor_extra_smem(LL.max_fs), c->stream>>>(
            S, LL.list_off + LL.n_tiny + LL.n_small, D.V.p, D.nV, D.Lv.p, D.nL, D.Uv.p, D.nU, D.Fbig.p, D.nFbig,
            D.status.p, D.uvec.p, D.nUvec, D.x.p, D.n);
      c->launches++;
    }
  }
  mark();  // factored
  // solves: small fronts by the warp kernels; per level the remaining fronts either all stage their
  // L panel in shared memory or none does
  auto rest_smem = [&](int lv, size_t& need, size_t& wneed, bool& fits, bool with_tiny = false) {
    const LevelLaunch& LL = D.levels[lv];
    need = wneed = 0;
    fits = true;
    for (int q = with_tiny ? 0 : LL.n_tiny; q < LL.n_tiny + LL.n_small + LL.n_big; q++) {
      int f = D.launch_list_host[LL.list_off + q];
      size_t fs = (size_t)D.sym.npiv[f] + D.sym.nupd[f];
      size_t nd = ((fs | 1) * D.sym.npiv[f] + fs + D.sym.npiv[f]) * sizeof(double);
      need = std::max(need, nd);
      wneed = std::max(wneed, (fs + D.sym.npiv[f]) * sizeof(double));
      if (nd > smem_limit) fits = false;
    }
  };
  // The forward sweep is fused into the factor kernels above.  SLAM_B200_SEPARATE_FORWARD=1 runs the
  // stand-alone forward kernels as well (same results; kept for A/B measurements).
  static const bool separate_forward = getenv("SLAM_B200_SEPARATE_FORWARD") != nullptr;
  for (int lv = 0; separate_forward && lv < nlv; lv++) {
    const LevelLaunch& LL = D.levels[lv];
    if (LL.n_tiny) {  // warp kernel: registers + shuffles, L read straight from global, no staging
      dim3 grid((LL.n_tiny + TINY_WARPS - 1) / TINY_WARPS, D.R);
      forward_tiny_kernel<<<grid, TINY_WARPS * 32, 0, c->stream>>>(S, LL.list_off, LL.n_tiny, D.V.p, D.nV, D.Lv.p,
                                                                  D.nL, D.uvec.p, D.nUvec, D.x.p, D.n);
      c->launches++;
    }
    const int nrest = LL.n_small + LL.n_big;
    if (!nrest) continue;
    dim3 grid(nrest, D.R);
    size_t need, wneed;
    bool fits;
    rest_smem(lv, need, wneed, fits);
    if (fits)
      forward_kernel<true><<<grid, SOLVE_THREADS, need, c->stream>>>(S, LL.list_off + LL.n_tiny, D.V.p, D.nV, D.Lv.p,
                                                                   D.nL, D.uvec.p, D.nUvec, D.x.p, D.n);
    else
      forward_kernel<false><<<grid, SOLVE_THREADS, wneed, c->stream>>>(S, LL.list_off + LL.n_tiny, D.V.p, D.nV,
                                                                     D.Lv.p, D.nL, D.uvec.p, D.nUvec, D.x.p, D.n);
    c->launches++;
  }
  mark();  // forward done
  for (int lv = nlv - 1; lv >= 0; lv--) {
    const LevelLaunch& LL = D.levels[lv];
    if (fuse_roots && lv == nlv - 1) continue;  // solved inside the factor kernel
    const int nrest = LL.n_small + LL.n_big;
    const bool merged = nrest && LL.n_big == 0 && merge_classes(LL);  // one launch for a mixed level, as above
    if (nrest) {
      dim3 grid(nrest + (merged ? LL.n_tiny : 0), D.R);
      const int off = LL.list_off + (merged ? 0 : LL.n_tiny);
      size_t need, wneed;
      bool fits;
      rest_smem(lv, need, wneed, fits, merged);
      if (fits)
        SLAM_CUDA_TRY(c, launch_front(backward_kernel<true>, grid, dim3(SOLVE_THREADS), need, c->stream, S, off, D.Lv.p, D.nL, D.x.p, D.n, early_l));
      else
        SLAM_CUDA_TRY(c, launch_front(backward_kernel<false>, grid, dim3(SOLVE_THREADS), wneed, c->stream, S, off, D.Lv.p, D.nL, D.x.p, D.n, early_l));
      early_l = 1;
      c->launches++;
    }
    if (LL.n_tiny && !merged) {
      dim3 grid((LL.n_tiny + TINY_WARPS - 1) / TINY_WARPS, D.R);
      SLAM_CUDA_TRY(c, launch_front(backward_tiny_kernel, grid, dim3(TINY_WARPS * 32), 0, c->stream, S, LL.list_off, LL.n_tiny, D.Lv.p, D.nL, D.x.p, D.n));
      early_l = 1;
      c->launches++;
    }
  }
  SLAM_CUDA_TRY(c, cudaGetLastError());
  mark();  // backward done
  int rc = graph_enqueue_update(c);
  mark();  // updated
  return rc;
}

// One Gauss-Newton iteration.  Kernel arguments do not change between iterations (the chi2 slot
// is a device-side counter), so the ~40 launches are captured once into a CUDA graph and replayed.
int graph_enqueue_iteration(slam_b200_ctx* c) {
  DeviceSystem& D = *c->sys;
  static const bool env_no_graph = getenv("SLAM_B200_NO_CUDA_GRAPH") != nullptr;
  // the legacy / per-thread default streams cannot be captured
  const bool no_graph = env_no_graph || c->stream == nullptr || c->stream == cudaStreamLegacy ||
                        c->stream == cudaStreamPerThread;
  if (int rc = solver_init_attrs(c)) return rc;
  if (D.profile || no_graph) {
    if (D.profile) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, c->stream);
      D.prof_events.push_back(e);
    }
    int rc = graph_enqueue_assemble(c, 0, D.P, false);
    if (rc) return rc;
    if (D.profile) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, c->stream);
      D.prof_events.push_back(e);
    }
    return graph_enqueue_solve(c);
  }
  if (!D.iter_graph) {
    NvtxRange nvtx_range("slam_b200/capture GN iteration graph");
    long before = c->launches;
    cudaGraph_t graph = nullptr;
    SLAM_CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = graph_enqueue_assemble(c, 0, D.P, false);
    if (!rc) rc = graph_enqueue_solve(c);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    D.launches_captured = c->launches - before;
    c->launches = before;
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    SLAM_CUDA_TRY(c, e);
    SLAM_CUDA_TRY(c, cudaGraphInstantiate(&D.iter_graph, graph, 0));
    cudaGraphDestroy(graph);
    // kernel nodes of one replay (every enqueue counts its launches): bookkeeping for slam_b200_launch_count
    D.launches_per_iter = (int)D.launches_captured;
  }
  SLAM_CUDA_TRY(c, cudaGraphLaunch(D.iter_graph, c->stream));
  c->launches += D.launches_per_iter;
  D.assembled = true;
  return 0;
}

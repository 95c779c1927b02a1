// graph.cu -- Gauss-Newton linearisation + normal-equation assembly + state update (sm_100a),
// and the host-side structure pass that prepares them.
//
// Replaces g2o's SparseOptimizer::initializeOptimization / buildIndexMapping (active set and
// Hessian index mapping), BlockSolver::buildStructure (block pattern), BlockSolver::buildSystem
// (EdgeSE2 / EdgeSE2PointXY computeError + linearizeOplus + BaseBinaryEdge::constructQuadraticForm)
// , SparseOptimizer::update (VertexSE2 / VertexPointXY oplusImpl) and computeActiveErrors /
// activeRobustChi2, which the reference reaches from Slam::optimizeGraph (slam.cpp:461-484).
//
// Assembly is "owner computes": one thread per pose walks that pose's edges (landmark edges of a
// pose are contiguous -- performSLAM inserts them per frame), accumulates the pose's diagonal block
// and rhs in registers and writes every off-diagonal block exactly once; one thread per landmark
// gathers its diagonal block and rhs over the edges that see it (recomputing the 2x2 part of the
// linearisation instead of exchanging it through memory).  No atomics, fixed summation order,
// bit-reproducible run to run.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>
#include <unordered_map>

#include <climits>

#include "graph_dev.h"

namespace {

__device__ __forceinline__ double normalize_theta_dev(double theta) {  // g2o stuff/misc.h
  const double pi = 3.14159265358979323846;
  if (theta >= -pi && theta < pi) return theta;
  double multiplier = floor(theta / (2 * pi));
  theta = theta - multiplier * 2 * pi;
  if (theta >= pi) theta -= 2 * pi;
  if (theta < -pi) theta += 2 * pi;
  return theta;
}

struct AsmArgs {
  int P, L, Eo, El;
  long estStride, measStride, nV;
  const double* est;
  const double* meas;
  double* V;
  const unsigned char* pose_free;
  const unsigned char* lm_free;
  const int* el_start;
  const int4* el_rec;  // {pose, landmark, block slot, flags} of every pose-sorted landmark edge
  const double* el_info;
  const int *lm_start, *lm_edges, *lmo_pose;
  const double* lmo_info;
  double* trig;  // [2][P] per replica: sin, cos of every pose heading, written by the pose kernel
  const int *eo_i, *eo_j, *eo_slot, *eo_flags, *po_start, *po_list;
  const double* eo_info;
  double* chi2_part;
  int chi2_blocks;
  double* chi2;
  int chi2_cap;
  const int* status;  // per replica [fail flag, iterations done]; the latter is the chi2 slot
};

constexpr int ASM_THREADS = 128;
constexpr int ASM_DEFAULT_VARIANT = 24;  // see graph_enqueue_assemble

// Pose-centred assembly, one WARP per 32 consecutive poses ("warp-aggregated block scatter"):
//  * the landmark edges of those poses are one contiguous range of the pose-sorted edge arrays, so
//    the lanes walk it 32 edges at a time with fully coalesced loads (each byte fetched once);
//  * every lane linearises one edge (pose trig comes from a per-warp shared-memory cache: sincos is
//    evaluated once per pose, not once per edge) and writes its off-diagonal block directly;
//  * the nine numbers an edge adds to its pose's diagonal block and rhs are combined by a segmented
//    shuffle scan over the lanes of the same pose; the tail lane of each segment adds the segment
//    sum to that pose's accumulator in shared memory (one writer per pose per step -> no atomics,
//    fixed order);
//  * finally lane l finishes pose l: its (at most a few) pose-pose edges, then one write of the
//    diagonal block and rhs.
constexpr int ASM_WARPS = ASM_THREADS / 32;

// minBlocksPerSM = 6 caps the kernel at 80 registers: the register-hungry part (3x3 algebra of the
// pose-pose edges) runs once per pose and may spill; the per-edge loop needs the occupancy to hide
// the latency of the landmark gather
//
// COALESCE: the blocks a warp step produces are contiguous in V (slots are handed out in edge
// order: 32 edges -> 192 consecutive doubles; 32 poses -> 96 + 288), but a lane holds one whole
// block, so direct stores are strided (every store instruction touches 12-18 lines).  The blocks
// are therefore transposed through shared memory and written with unit-stride stores whenever the
// step's slots are in fact contiguous (warp-uniform test; anything else -- duplicates, inactive or
// fixed vertices, transposed storage is fine -- falls back to the direct stores).
template <bool CHI2_ONLY, bool COALESCE>
__global__ void __launch_bounds__(ASM_THREADS, 6)
assemble_pose_kernel(AsmArgs a, int p0, int p1) {
  __shared__ double s_pose[ASM_WARPS][4][32];   // x, y, sin, cos of the warp's poses
  __shared__ double s_acc[ASM_WARPS][9][32];    // h00 h01 h02 h11 h12 h22 b0 b1 b2 per pose
  __shared__ double s_o[COALESCE ? ASM_WARPS : 1][COALESCE ? 288 : 1];  // one step's blocks, block-major
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.y;
  const double* est = a.est + (size_t)r * a.estStride;
  const double* meas = a.meas + (size_t)r * a.measStride;
  double* V = a.V + (size_t)r * a.nV;
  const int P = a.P, L = a.L, El = a.El, Eo = a.Eo;
  const int pw0 = p0 + (blockIdx.x * ASM_WARPS + wid) * 32;  // first pose of this warp
  const int p = pw0 + lane;
  double chi = 0;
  double px = 0, py = 0, pt = 0, s = 0, c = 1;
  bool free = false;
  if (p < p1) {
    px = est[p]; py = est[P + p]; pt = est[2 * P + p];
    sincos(pt, &s, &c);
    free = a.pose_free[p] != 0;
    if (!CHI2_ONLY) {  // the landmark kernel reuses the trig instead of recomputing it per edge
      double* trig = a.trig + (size_t)r * 2 * P;
      trig[p] = s;
      trig[P + p] = c;
    }
  }
  s_pose[wid][0][lane] = px; s_pose[wid][1][lane] = py; s_pose[wid][2][lane] = s; s_pose[wid][3][lane] = c;
#pragma unroll
  for (int k = 0; k < 9; k++) s_acc[wid][k][lane] = 0.0;
  __syncwarp();
  if (pw0 < p1) {  // warp-uniform
    const int pw1 = min(pw0 + 32, p1);
    const int e_begin = a.el_start[pw0], e_end = a.el_start[pw1];
    for (int base = e_begin; base < e_end; base += 32) {
      const int e = base + lane;
      const bool valid = e < e_end;
      int pl = -1 - lane;  // unique key for idle lanes
      double v[9];
#pragma unroll
      for (int k = 0; k < 9; k++) v[k] = 0.0;
      int fl = 0;
      double o[6];
      int slot = 0;
      bool wr = false;  // this lane stores o[] at V + slot
      if (valid) {
        fl = a.el_rec[e].w;
        pl = a.el_rec[e].x - pw0;
      }
      if (valid && (fl & EF_ACTIVE)) {
        const int l = a.el_rec[e].y;
        const double qx = s_pose[wid][0][pl], qy = s_pose[wid][1][pl], qs = s_pose[wid][2][pl], qc = s_pose[wid][3][pl];
        const double dx = est[3 * P + l] - qx, dy = est[3 * P + L + l] - qy;
        const double i00 = a.el_info[e], i01 = a.el_info[El + e], i11 = a.el_info[2 * El + e];
        const double j02 = -qs * dx + qc * dy;  // d e_x / d theta
        const double j12 = -qc * dx - qs * dy;  // d e_y / d theta
        const double ex = qc * dx + qs * dy - meas[e];
        const double ey = j02 - meas[El + e];
        chi += ex * (i00 * ex + i01 * ey) + ey * (i01 * ex + i11 * ey);
        if (!CHI2_ONLY && a.pose_free[pw0 + pl]) {
          // A = Ji^T Omega, Ji = [[-c, -s, j02], [s, -c, j12]]
          const double a00 = -qc * i00 + qs * i01, a01 = -qc * i01 + qs * i11;
          const double a10 = -qs * i00 - qc * i01, a11 = -qs * i01 - qc * i11;
          const double a20 = j02 * i00 + j12 * i01, a21 = j02 * i01 + j12 * i11;
          v[6] = -(a00 * ex + a01 * ey);
          v[7] = -(a10 * ex + a11 * ey);
          v[8] = -(a20 * ex + a21 * ey);
          v[0] = a00 * (-qc) + a01 * qs;
          v[1] = a00 * (-qs) + a01 * (-qc);
          v[2] = a00 * j02 + a01 * j12;
          v[3] = a10 * (-qs) + a11 * (-qc);
          v[4] = a10 * j02 + a11 * j12;
          v[5] = a20 * j02 + a21 * j12;
          if (fl & EF_OFFDIAG) {  // Ji^T Omega Jl, Jl = [[c, s], [-s, c]]
            const double B[6] = {a00 * qc - a01 * qs, a00 * qs + a01 * qc, a10 * qc - a11 * qs,
                                 a10 * qs + a11 * qc, a20 * qc - a21 * qs, a20 * qs + a21 * qc};
            slot = a.el_rec[e].z;
            if (fl & EF_TRANS) {  // stored landmark rows x pose columns (2x3)
              o[0] = B[0]; o[1] = B[2]; o[2] = B[4]; o[3] = B[1]; o[4] = B[3]; o[5] = B[5];
            } else {
#pragma unroll
              for (int k = 0; k < 6; k++) o[k] = B[k];
            }
            if (!COALESCE && (fl & EF_FIRST)) {
              double* hv = V + slot;
#pragma unroll
              for (int k = 0; k < 6; k++) hv[k] = o[k];
            }
            wr = (fl & EF_FIRST) != 0;
          }
        }
      }
      if (COALESCE && !CHI2_ONLY) {
        const int nvalid = min(32, e_end - base);
        const int slot0 = __shfl_sync(0xffffffffu, slot, 0);
        // every valid lane stores a first-of-its-pair block and the slots are consecutive
        const bool contig = __all_sync(0xffffffffu, !valid || (wr && slot == slot0 + 6 * lane));
        if (contig) {
#pragma unroll
          for (int k = 0; k < 6; k++) s_o[wid][6 * lane + k] = o[k];
          __syncwarp();
          double* hv = V + slot0;
#pragma unroll
          for (int j = 0; j < 6; j++) {
            const int q = j * 32 + lane;
            if (q < 6 * nvalid) hv[q] = s_o[wid][q];
          }
          __syncwarp();
        } else if (wr) {
          double* hv = V + slot;
#pragma unroll
          for (int k = 0; k < 6; k++) hv[k] = o[k];
        }
      }
      if (!CHI2_ONLY) {
        // duplicate edges of one (pose, landmark) pair (the doubled first-cone edge, repeated
        // matches) add to the block their first edge stored: rare, so serialised in edge order
        unsigned dup = __ballot_sync(0xffffffffu, valid && (fl & EF_ACTIVE) && (fl & EF_OFFDIAG) && !(fl & EF_FIRST) &&
                                                      a.pose_free[pw0 + max(pl, 0)]);
        while (dup) {
          __syncwarp();
          const int src = __ffs(dup) - 1;
          if (lane == src) {
            double* hv = V + slot;
#pragma unroll
            for (int k = 0; k < 6; k++) hv[k] += o[k];
          }
          dup &= dup - 1;
        }
        // segmented inclusive scan over lanes of the same pose (edges are sorted by pose)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int plo = __shfl_up_sync(0xffffffffu, pl, d);
          const bool take = lane >= d && plo == pl;
#pragma unroll
          for (int k = 0; k < 9; k++) {
            const double t = __shfl_up_sync(0xffffffffu, v[k], d);
            if (take) v[k] += t;
          }
        }
        const int pln = __shfl_down_sync(0xffffffffu, pl, 1);
        if (valid && (lane == 31 || pln != pl)) {  // tail of a segment: the only writer of this pose now
#pragma unroll
          for (int k = 0; k < 9; k++) s_acc[wid][k][pl] += v[k];
        }
        __syncwarp();
      }
    }
  }
  double fb[3] = {0, 0, 0}, fh[6] = {0, 0, 0, 0, 0, 0};  // finished pose rhs / block (COALESCE)
  if (p < p1) {
    double h00 = s_acc[wid][0][lane], h01 = s_acc[wid][1][lane], h02 = s_acc[wid][2][lane];
    double h11 = s_acc[wid][3][lane], h12 = s_acc[wid][4][lane], h22 = s_acc[wid][5][lane];
    double b0 = s_acc[wid][6][lane], b1 = s_acc[wid][7][lane], b2 = s_acc[wid][8][lane];
    // ---- pose-pose edges incident to this pose (EdgeSE2) ----
    const int q0 = a.po_start[p], q1 = a.po_start[p + 1];
    for (int q = q0; q < q1; q++) {
      const int ent = a.po_list[q];
      const int e = ent >> 1, side = ent & 1;
      const int fl = a.eo_flags[e];
      if (!(fl & EF_ACTIVE)) continue;
      const int i = a.eo_i[e], j = a.eo_j[e];
      double xi, yi, ti, xj, yj, tj, si, ci;
      if (side == 0) {
        xi = px; yi = py; ti = pt; si = s; ci = c;
        xj = est[j]; yj = est[P + j]; tj = est[2 * P + j];
      } else {
        xj = px; yj = py; tj = pt;
        xi = est[i]; yi = est[P + i]; ti = est[2 * P + i];
        sincos(ti, &si, &ci);
      }
      const double* mo = meas + 2 * (size_t)El;
      const double zx = mo[e], zy = mo[Eo + e], zt = mo[2 * Eo + e];
      const double o00 = a.eo_info[e], o01 = a.eo_info[Eo + e], o02 = a.eo_info[2 * Eo + e],
                   o11 = a.eo_info[3 * Eo + e], o12 = a.eo_info[4 * Eo + e], o22 = a.eo_info[5 * Eo + e];
      const double dtx = xj - xi, dty = yj - yi;
      const double tx = ci * dtx + si * dty, ty = -si * dtx + ci * dty;  // R_i^T (t_j - t_i)
      double sz, cz;
      sincos(zt, &sz, &cz);
      double err[3];
      err[0] = cz * (tx - zx) + sz * (ty - zy);
      err[1] = -sz * (tx - zx) + cz * (ty - zy);
      err[2] = normalize_theta_dev(tj - ti - zt);
      const double Om[3][3] = {{o00, o01, o02}, {o01, o11, o12}, {o02, o12, o22}};
      double Oe[3];
#pragma unroll
      for (int k = 0; k < 3; k++) Oe[k] = Om[k][0] * err[0] + Om[k][1] * err[1] + Om[k][2] * err[2];
      if (side == 0) chi += err[0] * Oe[0] + err[1] * Oe[1] + err[2] * Oe[2];
      if (CHI2_ONLY) continue;
      // Ji = Z A, Jj = Z B with Z = diag(R(-zt), 1)
      const double A0[3] = {-ci, -si, ty}, A1[3] = {si, -ci, -tx};
      const double B0[3] = {ci, si, 0}, B1[3] = {-si, ci, 0};
      double Ji[3][3], Jj[3][3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        Ji[0][k] = cz * A0[k] + sz * A1[k];
        Ji[1][k] = -sz * A0[k] + cz * A1[k];
        Jj[0][k] = cz * B0[k] + sz * B1[k];
        Jj[1][k] = -sz * B0[k] + cz * B1[k];
      }
      Ji[2][0] = 0; Ji[2][1] = 0; Ji[2][2] = -1;
      Jj[2][0] = 0; Jj[2][1] = 0; Jj[2][2] = 1;
      if (free) {
        double AtO[3][3];  // Jm^T Omega
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int m = 0; m < 3; m++) {
            double jm0 = side == 0 ? Ji[0][k] : Jj[0][k];
            double jm1 = side == 0 ? Ji[1][k] : Jj[1][k];
            double jm2 = side == 0 ? Ji[2][k] : Jj[2][k];
            AtO[k][m] = jm0 * Om[0][m] + jm1 * Om[1][m] + jm2 * Om[2][m];
          }
        b0 -= AtO[0][0] * err[0] + AtO[0][1] * err[1] + AtO[0][2] * err[2];
        b1 -= AtO[1][0] * err[0] + AtO[1][1] * err[1] + AtO[1][2] * err[2];
        b2 -= AtO[2][0] * err[0] + AtO[2][1] * err[1] + AtO[2][2] * err[2];
        double Hm[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int m = k; m < 3; m++) {
            double v = 0;
#pragma unroll
            for (int t = 0; t < 3; t++) v += AtO[k][t] * (side == 0 ? Ji[t][m] : Jj[t][m]);
            Hm[k][m] = v;
          }
        h00 += Hm[0][0]; h01 += Hm[0][1]; h02 += Hm[0][2];
        h11 += Hm[1][1]; h12 += Hm[1][2]; h22 += Hm[2][2];
      }
      if ((fl & EF_OFFDIAG) && p == min(i, j)) {  // Ji^T Omega Jj, written by one owner thread
        double AtO[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int m = 0; m < 3; m++)
            AtO[k][m] = Ji[0][k] * Om[0][m] + Ji[1][k] * Om[1][m] + Ji[2][k] * Om[2][m];
        double* hv = V + a.eo_slot[e];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int m = 0; m < 3; m++) {
            double v = AtO[k][0] * Jj[0][m] + AtO[k][1] * Jj[1][m] + AtO[k][2] * Jj[2][m];
            int o = (fl & EF_TRANS) ? (m * 3 + k) : (k * 3 + m);
            if (fl & EF_FIRST) hv[o] = v;
            else hv[o] += v;
          }
      }
    }
    if (!CHI2_ONLY && !COALESCE && free) {
      double* bp = V + 6 * (size_t)L + 3 * (size_t)p;
      bp[0] = b0; bp[1] = b1; bp[2] = b2;
      double* hp = V + 6 * (size_t)L + 3 * (size_t)P + 9 * (size_t)p;
      hp[0] = h00; hp[1] = h01; hp[2] = h02;
      hp[3] = h01; hp[4] = h11; hp[5] = h12;
      hp[6] = h02; hp[7] = h12; hp[8] = h22;
    }
    if (COALESCE) { fb[0] = b0; fb[1] = b1; fb[2] = b2; fh[0] = h00; fh[1] = h01; fh[2] = h02; fh[3] = h11; fh[4] = h12; fh[5] = h22; }
  }
  if (COALESCE && !CHI2_ONLY && pw0 < p1) {  // warp-uniform
    const int np = min(32, p1 - pw0);
    const bool allfree = __all_sync(0xffffffffu, p >= p1 || free);
    if (allfree) {
      double* bp = V + 6 * (size_t)L + 3 * (size_t)pw0;
      double* hp = V + 6 * (size_t)L + 3 * (size_t)P + 9 * (size_t)pw0;
      s_o[wid][3 * lane] = fb[0]; s_o[wid][3 * lane + 1] = fb[1]; s_o[wid][3 * lane + 2] = fb[2];
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int q = j * 32 + lane;
        if (q < 3 * np) bp[q] = s_o[wid][q];
      }
      __syncwarp();
      double* so = &s_o[wid][9 * lane];
      so[0] = fh[0]; so[1] = fh[1]; so[2] = fh[2];
      so[3] = fh[1]; so[4] = fh[3]; so[5] = fh[4];
      so[6] = fh[2]; so[7] = fh[4]; so[8] = fh[5];
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 9; j++) {
        const int q = j * 32 + lane;
        if (q < 9 * np) hp[q] = s_o[wid][q];
      }
    } else if (p < p1 && free) {
      double* bp = V + 6 * (size_t)L + 3 * (size_t)p;
      bp[0] = fb[0]; bp[1] = fb[1]; bp[2] = fb[2];
      double* hp = V + 6 * (size_t)L + 3 * (size_t)P + 9 * (size_t)p;
      hp[0] = fh[0]; hp[1] = fh[1]; hp[2] = fh[2];
      hp[3] = fh[1]; hp[4] = fh[3]; hp[5] = fh[4];
      hp[6] = fh[2]; hp[7] = fh[4]; hp[8] = fh[5];
    }
  }
  // chi2: fixed-order block reduction -> one partial per block
  __shared__ double red[ASM_THREADS / 32];
  for (int o = 16; o; o >>= 1) chi += __shfl_down_sync(0xffffffffu, chi, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = chi;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < ASM_THREADS / 32; w++) t += red[w];
    a.chi2_part[(size_t)r * a.chi2_blocks + blockIdx.x] = t;
  }
}

// ---- pipelined variant of the pose-centred assembly -------------------------------------------------
// Same decomposition and the same arithmetic, operation for operation, as assemble_pose_kernel
// <false, true> (so V is bit-identical), but the loads are restructured.  ncu on the kernel above:
// 66 % of warp time is long_scoreboard spread over FOUR dependent load sites per 32-edge step
// (flags -> landmark index / information / measurement -> landmark estimate -> slot), at 24 warps/SM;
// no resource is saturated, the step is a chain of exposed round trips.  Here
//  * the four ints of an edge are one 16-byte record (one load instead of four, and no load sits
//    behind a branch on the flags of another);
//  * the pose_free test of the edge loop is a ballot mask in a register;
//  * DEPTH 1: the record + payload of step k+1 are loaded at the top of step k and the landmark
//    estimate of step k+1 is gathered at its bottom: one exposed round trip per step instead of four;
//    DEPTH 2: records run two steps ahead, so the gather of step k+1 is issued at the top of step k
//    too and no load of the steady state is exposed;
//  * the pose-pose epilogue loads the incidence entries of a pose two at a time (a trackdrive pose
//    has exactly two: its odometry edges) with the dependent index loads of both in flight together,
//    and reads the sin/cos of a neighbour inside the warp's pose range from the shared-memory cache.
struct ElPay { double i00, i01, i11, m0, m1; };

template <int DEPTH, int MINB>
__global__ void __launch_bounds__(ASM_THREADS, MINB)
assemble_pose_pipe_kernel(AsmArgs a, int p0, int p1) {
  __shared__ double s_pose[ASM_WARPS][4][32];
  __shared__ double s_acc[ASM_WARPS][9][32];
  __shared__ double s_o[ASM_WARPS][288];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = blockIdx.y;
  const double* __restrict__ est = a.est + (size_t)r * a.estStride;
  const double* __restrict__ meas = a.meas + (size_t)r * a.measStride;
  double* V = a.V + (size_t)r * a.nV;
  const int P = a.P, L = a.L, El = a.El, Eo = a.Eo;
  const int pw0 = p0 + (blockIdx.x * ASM_WARPS + wid) * 32;
  const int p = pw0 + lane;
  double chi = 0;
  double px = 0, py = 0, pt = 0, s = 0, c = 1;
  bool free = false;
  int q0 = 0, q1 = 0;
  if (p < p1) {
    px = est[p]; py = est[P + p]; pt = est[2 * P + p];
    free = a.pose_free[p] != 0;
    q0 = a.po_start[p]; q1 = a.po_start[p + 1];
    sincos(pt, &s, &c);
    double* trig = a.trig + (size_t)r * 2 * P;
    trig[p] = s;
    trig[P + p] = c;
  }
  const unsigned freemask = __ballot_sync(0xffffffffu, free);
  s_pose[wid][0][lane] = px; s_pose[wid][1][lane] = py; s_pose[wid][2][lane] = s; s_pose[wid][3][lane] = c;
#pragma unroll
  for (int k = 0; k < 9; k++) s_acc[wid][k][lane] = 0.0;
  __syncwarp();
  if (pw0 < p1) {  // warp-uniform
    const int pw1 = min(pw0 + 32, p1);
    const int e_begin = a.el_start[pw0], e_end = a.el_start[pw1];
    const double* __restrict__ inf0 = a.el_info;
    const double* __restrict__ inf1 = a.el_info + El;
    const double* __restrict__ inf2 = a.el_info + 2 * (size_t)El;
    const double* __restrict__ elx = est + 3 * (size_t)P;
    const double* __restrict__ ely = est + 3 * (size_t)P + L;
    auto ld_ix = [&](int e) -> int4 {  // idle lanes: unique pose key, flags 0
      return e < e_end ? __ldg(a.el_rec + e) : make_int4(pw0 - 1 - lane, 0, 0, 0);
    };
    auto ld_pay = [&](int e) -> ElPay {
      ElPay y = {0, 0, 0, 0, 0};
      if (e < e_end) {
        y.i00 = __ldg(inf0 + e); y.i01 = __ldg(inf1 + e); y.i11 = __ldg(inf2 + e);
        y.m0 = __ldg(meas + e); y.m1 = __ldg(meas + El + e);
      }
      return y;
    };
    auto ld_est = [&](const int4& ix) -> double2 {
      double2 v = make_double2(0, 0);
      if (ix.w & EF_ACTIVE) { v.x = elx[ix.y]; v.y = ely[ix.y]; }
      return v;
    };
    int4 ixA = ld_ix(e_begin + lane);
    int4 ixB = make_int4(0, 0, 0, 0);
    ElPay payA = ld_pay(e_begin + lane);
    if (DEPTH >= 2) ixB = ld_ix(e_begin + 32 + lane);
    double2 estA = ld_est(ixA);
    for (int base = e_begin; base < e_end; base += 32) {
      // ---- loads of the coming steps, issued before anything of this step is used ----
      int4 ixC = make_int4(0, 0, 0, 0);
      ElPay payB = {0, 0, 0, 0, 0};
      double2 estB = make_double2(0, 0);
      if (DEPTH == 1) { ixB = ld_ix(base + 32 + lane); payB = ld_pay(base + 32 + lane); }
      if (DEPTH >= 2) { ixC = ld_ix(base + 64 + lane); payB = ld_pay(base + 32 + lane); estB = ld_est(ixB); }
      // ---- this step ----
      const bool valid = base + lane < e_end;
      const int fl = ixA.w;
      const int pl = ixA.x - pw0;
      double v[9];
#pragma unroll
      for (int k = 0; k < 9; k++) v[k] = 0.0;
      double o[6];
      int slot = 0;
      bool wr = false;
      const bool pfree = valid && ((freemask >> (pl & 31)) & 1u);
      if (valid && (fl & EF_ACTIVE)) {
        const double qx = s_pose[wid][0][pl], qy = s_pose[wid][1][pl], qs = s_pose[wid][2][pl], qc = s_pose[wid][3][pl];
        const double dx = estA.x - qx, dy = estA.y - qy;
        const double i00 = payA.i00, i01 = payA.i01, i11 = payA.i11;
        const double j02 = -qs * dx + qc * dy;
        const double j12 = -qc * dx - qs * dy;
        const double ex = qc * dx + qs * dy - payA.m0;
        const double ey = j02 - payA.m1;
        chi += ex * (i00 * ex + i01 * ey) + ey * (i01 * ex + i11 * ey);
        if (pfree) {
          const double a00 = -qc * i00 + qs * i01, a01 = -qc * i01 + qs * i11;
          const double a10 = -qs * i00 - qc * i01, a11 = -qs * i01 - qc * i11;
          const double a20 = j02 * i00 + j12 * i01, a21 = j02 * i01 + j12 * i11;
          v[6] = -(a00 * ex + a01 * ey);
          v[7] = -(a10 * ex + a11 * ey);
          v[8] = -(a20 * ex + a21 * ey);
          v[0] = a00 * (-qc) + a01 * qs;
          v[1] = a00 * (-qs) + a01 * (-qc);
          v[2] = a00 * j02 + a01 * j12;
          v[3] = a10 * (-qs) + a11 * (-qc);
          v[4] = a10 * j02 + a11 * j12;
          v[5] = a20 * j02 + a21 * j12;
          if (fl & EF_OFFDIAG) {
            const double B[6] = {a00 * qc - a01 * qs, a00 * qs + a01 * qc, a10 * qc - a11 * qs,
                                 a10 * qs + a11 * qc, a20 * qc - a21 * qs, a20 * qs + a21 * qc};
            slot = ixA.z;
            if (fl & EF_TRANS) {
              o[0] = B[0]; o[1] = B[2]; o[2] = B[4]; o[3] = B[1]; o[4] = B[3]; o[5] = B[5];
            } else {
#pragma unroll
              for (int k = 0; k < 6; k++) o[k] = B[k];
            }
            wr = (fl & EF_FIRST) != 0;
          }
        }
      }
      {
        const int nvalid = min(32, e_end - base);
        const int slot0 = __shfl_sync(0xffffffffu, slot, 0);
        const bool contig = __all_sync(0xffffffffu, !valid || (wr && slot == slot0 + 6 * lane));
        if (contig) {
#pragma unroll
          for (int k = 0; k < 6; k++) s_o[wid][6 * lane + k] = o[k];
          __syncwarp();
          double* hv = V + slot0;
#pragma unroll
          for (int j = 0; j < 6; j++) {
            const int q = j * 32 + lane;
            if (q < 6 * nvalid) hv[q] = s_o[wid][q];
          }
          __syncwarp();
        } else if (wr) {
          double* hv = V + slot;
#pragma unroll
          for (int k = 0; k < 6; k++) hv[k] = o[k];
        }
      }
      unsigned dup = __ballot_sync(0xffffffffu, valid && (fl & EF_ACTIVE) && (fl & EF_OFFDIAG) && !(fl & EF_FIRST) && pfree);
      while (dup) {
        __syncwarp();
        const int src = __ffs(dup) - 1;
        if (lane == src) {
          double* hv = V + slot;
#pragma unroll
          for (int k = 0; k < 6; k++) hv[k] += o[k];
        }
        dup &= dup - 1;
      }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int plo = __shfl_up_sync(0xffffffffu, pl, d);
        const bool take = lane >= d && plo == pl;
#pragma unroll
        for (int k = 0; k < 9; k++) {
          const double t = __shfl_up_sync(0xffffffffu, v[k], d);
          if (take) v[k] += t;
        }
      }
      const int pln = __shfl_down_sync(0xffffffffu, pl, 1);
      if (valid && (lane == 31 || pln != pl)) {
#pragma unroll
        for (int k = 0; k < 9; k++) s_acc[wid][k][pl] += v[k];
      }
      __syncwarp();
      // ---- rotate the pipeline ----
      if (DEPTH == 0) { ixA = ld_ix(base + 32 + lane); payA = ld_pay(base + 32 + lane); estA = ld_est(ixA); }
      if (DEPTH == 1) { ixA = ixB; payA = payB; estA = ld_est(ixA); }
      if (DEPTH >= 2) { ixA = ixB; ixB = ixC; payA = payB; estA = estB; }
    }
  }
  double fb[3] = {0, 0, 0}, fh[6] = {0, 0, 0, 0, 0, 0};
  if (p < p1) {
    double h00 = s_acc[wid][0][lane], h01 = s_acc[wid][1][lane], h02 = s_acc[wid][2][lane];
    double h11 = s_acc[wid][3][lane], h12 = s_acc[wid][4][lane], h22 = s_acc[wid][5][lane];
    double b0 = s_acc[wid][6][lane], b1 = s_acc[wid][7][lane], b2 = s_acc[wid][8][lane];
    const double* __restrict__ mo = meas + 2 * (size_t)El;
    // one pose-pose edge; (xo, yo, to) is the OTHER pose of the edge
    auto po_edge = [&](int e, int side, int fl, int i, int j, int eslot, double xo, double yo, double to) {
      double xi, yi, ti, xj, yj, tj, si, ci;
      if (side == 0) {
        xi = px; yi = py; ti = pt; si = s; ci = c;
        xj = xo; yj = yo; tj = to;
      } else {
        xj = px; yj = py; tj = pt;
        xi = xo; yi = yo; ti = to;
        const int il = i - pw0;
        if (il >= 0 && il < 32 && i < p1) { si = s_pose[wid][2][il]; ci = s_pose[wid][3][il]; }  // same sincos(ti), cached
        else sincos(ti, &si, &ci);
      }
      const double zx = mo[e], zy = mo[Eo + e], zt = mo[2 * Eo + e];
      const double o00 = a.eo_info[e], o01 = a.eo_info[Eo + e], o02 = a.eo_info[2 * Eo + e],
                   o11 = a.eo_info[3 * Eo + e], o12 = a.eo_info[4 * Eo + e], o22 = a.eo_info[5 * Eo + e];
      const double dtx = xj - xi, dty = yj - yi;
      const double tx = ci * dtx + si * dty, ty = -si * dtx + ci * dty;
      double sz, cz;
      sincos(zt, &sz, &cz);
      double err[3];
      err[0] = cz * (tx - zx) + sz * (ty - zy);
      err[1] = -sz * (tx - zx) + cz * (ty - zy);
      err[2] = normalize_theta_dev(tj - ti - zt);
      const double Om[3][3] = {{o00, o01, o02}, {o01, o11, o12}, {o02, o12, o22}};
      double Oe[3];
#pragma unroll
      for (int k = 0; k < 3; k++) Oe[k] = Om[k][0] * err[0] + Om[k][1] * err[1] + Om[k][2] * err[2];
      if (side == 0) chi += err[0] * Oe[0] + err[1] * Oe[1] + err[2] * Oe[2];
      const double A0[3] = {-ci, -si, ty}, A1[3] = {si, -ci, -tx};
      const double B0[3] = {ci, si, 0}, B1[3] = {-si, ci, 0};
      double Ji[3][3], Jj[3][3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        Ji[0][k] = cz * A0[k] + sz * A1[k];
        Ji[1][k] = -sz * A0[k] + cz * A1[k];
        Jj[0][k] = cz * B0[k] + sz * B1[k];
        Jj[1][k] = -sz * B0[k] + cz * B1[k];
      }
      Ji[2][0] = 0; Ji[2][1] = 0; Ji[2][2] = -1;
      Jj[2][0] = 0; Jj[2][1] = 0; Jj[2][2] = 1;
      if (free) {
        double AtO[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int m = 0; m < 3; m++) {
            double jm0 = side == 0 ? Ji[0][k] : Jj[0][k];
            double jm1 = side == 0 ? Ji[1][k] : Jj[1][k];
            double jm2 = side == 0 ? Ji[2][k] : Jj[2][k];
            AtO[k][m] = jm0 * Om[0][m] + jm1 * Om[1][m] + jm2 * Om[2][m];
          }
        b0 -= AtO[0][0] * err[0] + AtO[0][1] * err[1] + AtO[0][2] * err[2];
        b1 -= AtO[1][0] * err[0] + AtO[1][1] * err[1] + AtO[1][2] * err[2];
        b2 -= AtO[2][0] * err[0] + AtO[2][1] * err[1] + AtO[2][2] * err[2];
        double Hm[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int m = k; m < 3; m++) {
            double vv = 0;
#pragma unroll
            for (int t = 0; t < 3; t++) vv += AtO[k][t] * (side == 0 ? Ji[t][m] : Jj[t][m]);
            Hm[k][m] = vv;
          }
        h00 += Hm[0][0]; h01 += Hm[0][1]; h02 += Hm[0][2];
        h11 += Hm[1][1]; h12 += Hm[1][2]; h22 += Hm[2][2];
      }
      if ((fl & EF_OFFDIAG) && p == min(i, j)) {
        double AtO[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int m = 0; m < 3; m++)
            AtO[k][m] = Ji[0][k] * Om[0][m] + Ji[1][k] * Om[1][m] + Ji[2][k] * Om[2][m];
        double* hv = V + eslot;
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int m = 0; m < 3; m++) {
            double vv = AtO[k][0] * Jj[0][m] + AtO[k][1] * Jj[1][m] + AtO[k][2] * Jj[2][m];
            int oo = (fl & EF_TRANS) ? (m * 3 + k) : (k * 3 + m);
            if (fl & EF_FIRST) hv[oo] = vv;
            else hv[oo] += vv;
          }
      }
    };
    for (int q = q0; q < q1; q += 2) {
      const bool two = q + 1 < q1;
      const int entA = a.po_list[q], entB = two ? a.po_list[q + 1] : 0;
      const int eA = entA >> 1, sideA = entA & 1, eB = entB >> 1, sideB = entB & 1;
      const int flA = a.eo_flags[eA], iA = a.eo_i[eA], jA = a.eo_j[eA], slA = a.eo_slot[eA];
      int flB = 0, iB = 0, jB = 0, slB = 0;
      if (two) { flB = a.eo_flags[eB]; iB = a.eo_i[eB]; jB = a.eo_j[eB]; slB = a.eo_slot[eB]; }
      const int oA = sideA == 0 ? jA : iA, oB = sideB == 0 ? jB : iB;
      const double xA = est[oA], yA = est[P + oA], tA = est[2 * P + oA];
      double xB = 0, yB = 0, tB = 0;
      if (two) { xB = est[oB]; yB = est[P + oB]; tB = est[2 * P + oB]; }
      if (flA & EF_ACTIVE) po_edge(eA, sideA, flA, iA, jA, slA, xA, yA, tA);
      if (two && (flB & EF_ACTIVE)) po_edge(eB, sideB, flB, iB, jB, slB, xB, yB, tB);
    }
    fb[0] = b0; fb[1] = b1; fb[2] = b2; fh[0] = h00; fh[1] = h01; fh[2] = h02; fh[3] = h11; fh[4] = h12; fh[5] = h22;
  }
  if (pw0 < p1) {  // warp-uniform
    const int np = min(32, p1 - pw0);
    const bool allfree = __all_sync(0xffffffffu, p >= p1 || free);
    if (allfree) {
      double* bp = V + 6 * (size_t)L + 3 * (size_t)pw0;
      double* hp = V + 6 * (size_t)L + 3 * (size_t)P + 9 * (size_t)pw0;
      s_o[wid][3 * lane] = fb[0]; s_o[wid][3 * lane + 1] = fb[1]; s_o[wid][3 * lane + 2] = fb[2];
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int q = j * 32 + lane;
        if (q < 3 * np) bp[q] = s_o[wid][q];
      }
      __syncwarp();
      double* so = &s_o[wid][9 * lane];
      so[0] = fh[0]; so[1] = fh[1]; so[2] = fh[2];
      so[3] = fh[1]; so[4] = fh[3]; so[5] = fh[4];
      so[6] = fh[2]; so[7] = fh[4]; so[8] = fh[5];
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 9; j++) {
        const int q = j * 32 + lane;
        if (q < 9 * np) hp[q] = s_o[wid][q];
      }
    } else if (p < p1 && free) {
      double* bp = V + 6 * (size_t)L + 3 * (size_t)p;
      bp[0] = fb[0]; bp[1] = fb[1]; bp[2] = fb[2];
      double* hp = V + 6 * (size_t)L + 3 * (size_t)P + 9 * (size_t)p;
      hp[0] = fh[0]; hp[1] = fh[1]; hp[2] = fh[2];
      hp[3] = fh[1]; hp[4] = fh[3]; hp[5] = fh[4];
      hp[6] = fh[2]; hp[7] = fh[4]; hp[8] = fh[5];
    }
  }
  __shared__ double red[ASM_THREADS / 32];
  for (int o = 16; o; o >>= 1) chi += __shfl_down_sync(0xffffffffu, chi, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = chi;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < ASM_THREADS / 32; w++) t += red[w];
    a.chi2_part[(size_t)r * a.chi2_blocks + blockIdx.x] = t;
  }
}

// one WARP per landmark: diagonal block + rhs over the edges (with pose in [p0,p1)) that see it.
// A landmark is seen from tens to hundreds of poses (24 per lap on the trackdrive), so the edge list
// is strided over the lanes and the five partial sums are combined with a fixed xor-shuffle tree
// (deterministic).  The 2x2 part of the linearisation is recomputed from the pose instead of being
// exchanged through memory.
constexpr int LM_PER_BLOCK = ASM_THREADS / 32;

// Peer exchange of the landmark part between the pose-range shards of ONE graph (config 5, one
// process per GPU).  Every rank owns a region of device memory opened by all the others through
// CUDA IPC: [world 64-bit flags | 2 parities x 6*cap doubles].  PEER mode of the landmark kernel
// writes the partial block + rhs of every landmark of this shard's range into the rank's OWN slot
// (b: 2 per landmark, then H: 4 per landmark) instead of V.  lm_exchange_sum_kernel, next on the
// stream, publishes `epoch` in every rank's flag word (the only remote stores), waits for all flags
// and PULLS: V[landmark part] = sum over ranks, ascending, of their slots, read straight
// from the peers' memory over NVLink with unit-stride 16-byte loads (deterministic, unlike a ring
// all-reduce; pushing 48-byte records from the landmark kernel's single writer lanes was measured
// 2x slower than NCCL -- small remote stores do not fill NVLink packets).  Two parities: a rank can
// start assembly k+1 while a slower rank still pulls assembly k from it; it cannot reach k+2 before
// every rank has raised flag k+1, i.e. finished pulling k.
constexpr int XCHG_HEADER = 1024;  // bytes reserved for the flags
struct PeerArgs {
  char* const* peer_tab;  // region base of every rank (device array)
  int world, rank, cap, parity;
  unsigned long long epoch;
};
__device__ __forceinline__ double* xchg_slot(char* base, int cap, int parity) {
  return reinterpret_cast<double*>(base + XCHG_HEADER) + (size_t)parity * 6 * (size_t)cap;
}

// G lanes per landmark (32, 16 or 8; 32/G landmarks per warp): with one warp per landmark the kernel
// executed 437 warp instructions per landmark on the 1M-pose graph (36 observers per landmark: 28 of
// the 64 lane slots of its two steps idle, a 5-value x 5-level shuffle tree per landmark) and was half
// issue-bound (ncu: issue slots 50 % busy at 40 % occupancy).  Smaller groups fill the lanes and
// shorten the tree; G is chosen from the mean number of observers per landmark (host, below).
template <bool CHI2_ONLY, bool UNROLL2, bool PEER, int G>
__global__ void __launch_bounds__(ASM_THREADS)
assemble_landmark_kernel(AsmArgs a, int p0, int p1, int chi2_nblocks, int l_first, int l_end, PeerArgs px) {
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  constexpr int GPW = 32 / G;  // landmarks per warp
  const int l = l_first + (blockIdx.x * LM_PER_BLOCK + (threadIdx.x >> 5)) * GPW + lane / G;
  const int r = blockIdx.y;
  const double* est = a.est + (size_t)r * a.estStride;
  const double* meas = a.meas + (size_t)r * a.measStride;
  double* V = a.V + (size_t)r * a.nV;
  const int P = a.P, L = a.L, El = a.El;
  const bool inrange = !CHI2_ONLY && l < l_end;
  const bool active = inrange && a.lm_free[l];
  if (PEER && inrange && !active && gl == 0) {
    // fixed / inactive landmark inside the range: its slot entries must read as zero
    double* slot = xchg_slot(px.peer_tab[px.rank], px.cap, px.parity);
    const int k = l - l_first;
    slot[2 * k] = 0.0; slot[2 * k + 1] = 0.0;
    double* hs = slot + 2 * (size_t)px.cap + 4 * (size_t)k;
    hs[0] = 0.0; hs[1] = 0.0; hs[2] = 0.0; hs[3] = 0.0;
  }
  if (!CHI2_ONLY) {  // all lanes stay together (the reduction below shuffles across the warp)
    double lx = 0, ly = 0;
    double h00 = 0, h01 = 0, h11 = 0, b0 = 0, b1 = 0;
    int q0 = 0, q1 = 0;
    if (active) {
      lx = est[3 * P + l]; ly = est[3 * P + L + l];
      q0 = a.lm_start[l]; q1 = a.lm_start[l + 1];
      if (p0 > 0 || p1 < P) {
        // pose-range shard: landmark edges are sorted by pose, so the shard's edges are the sorted
        // positions [el_start[p0], el_start[p1]); the landmark's list is ascending -> two bisections
        const int elo = a.el_start[p0], ehi = a.el_start[p1];
        int lo = q0, hi = q1;
        while (lo < hi) { int m = (lo + hi) >> 1; if (a.lm_edges[m] < elo) lo = m + 1; else hi = m; }
        const int b0_ = lo;
        hi = q1;
        while (lo < hi) { int m = (lo + hi) >> 1; if (a.lm_edges[m] < ehi) lo = m + 1; else hi = m; }
        q0 = b0_;
        q1 = lo;
      }
    }
    const double* trig = a.trig + (size_t)r * 2 * P;
    const double* mlm = meas + 2 * (size_t)El + 3 * (size_t)a.Eo;  // measurements in landmark order
    // everything indexed by q is laid out in landmark order: coalesced across the lanes.  Two
    // G-edge steps are loaded together (index -> gathers -> payload for both) before either is
    // evaluated, so the two dependent round trips of a step overlap with those of the next.
    const double* io0 = a.lmo_info;
    const double* io1 = a.lmo_info + El;
    const double* io2 = a.lmo_info + 2 * (size_t)El;
    const double* m0 = mlm;
    const double* m1 = mlm + El;
    const double* ex_ = est;
    const double* ey_ = est + P;
    const double* tc_ = trig + P;
    for (int qb = q0 + gl; qb < q1; qb += (UNROLL2 ? 2 * G : G)) {
      const int qa = qb, qc = qb + G;
      const bool vc = UNROLL2 && qc < q1;
      const int pa = __ldg(a.lmo_pose + qa);
      const int pc = vc ? __ldg(a.lmo_pose + qc) : -1;
      const bool aa = pa >= 0, ac = pc >= 0;  // p < 0: inactive edge
      double sa = 0, ca = 0, xa = 0, ya = 0, sc = 0, cc = 0, xc = 0, yc = 0;
      if (aa) { sa = trig[pa]; ca = tc_[pa]; xa = ex_[pa]; ya = ey_[pa]; }
      if (ac) { sc = trig[pc]; cc = tc_[pc]; xc = ex_[pc]; yc = ey_[pc]; }
      double za0 = 0, za1 = 0, ia0 = 0, ia1 = 0, ia2 = 0, zc0 = 0, zc1 = 0, ic0 = 0, ic1 = 0, ic2 = 0;
      if (aa) { za0 = __ldg(m0 + qa); za1 = __ldg(m1 + qa); ia0 = __ldg(io0 + qa); ia1 = __ldg(io1 + qa); ia2 = __ldg(io2 + qa); }
      if (ac) { zc0 = __ldg(m0 + qc); zc1 = __ldg(m1 + qc); ic0 = __ldg(io0 + qc); ic1 = __ldg(io1 + qc); ic2 = __ldg(io2 + qc); }
      auto accumulate = [&](double s, double c, double px_, double py_, double z0, double z1, double i00, double i01,
                            double i11) {
        const double dx = lx - px_, dy = ly - py_;
        const double ex = c * dx + s * dy - z0;
        const double ey = -s * dx + c * dy - z1;
        // A = Jl^T Omega, Jl = [[c, s], [-s, c]]
        const double a00 = c * i00 - s * i01, a01 = c * i01 - s * i11;
        const double a10 = s * i00 + c * i01, a11 = s * i01 + c * i11;
        b0 -= a00 * ex + a01 * ey;
        b1 -= a10 * ex + a11 * ey;
        h00 += a00 * c - a01 * s;
        h01 += a00 * s + a01 * c;
        h11 += a10 * s + a11 * c;
      };
      if (aa) accumulate(sa, ca, xa, ya, za0, za1, ia0, ia1, ia2);
      if (ac) accumulate(sc, cc, xc, yc, zc0, zc1, ic0, ic1, ic2);
    }
    __syncwarp();
#pragma unroll
    for (int o = G / 2; o; o >>= 1) {  // fixed xor tree inside the group of G lanes
      h00 += __shfl_xor_sync(0xffffffffu, h00, o);
      h01 += __shfl_xor_sync(0xffffffffu, h01, o);
      h11 += __shfl_xor_sync(0xffffffffu, h11, o);
      b0 += __shfl_xor_sync(0xffffffffu, b0, o);
      b1 += __shfl_xor_sync(0xffffffffu, b1, o);
    }
    if (active && gl == 0) {
      if (PEER) {
        double* slot = xchg_slot(px.peer_tab[px.rank], px.cap, px.parity);
        const int k = l - l_first;
        *reinterpret_cast<double2*>(slot + 2 * (size_t)k) = make_double2(b0, b1);
        double2* hs = reinterpret_cast<double2*>(slot + 2 * (size_t)px.cap + 4 * (size_t)k);
        hs[0] = make_double2(h00, h01);
        hs[1] = make_double2(h01, h11);
      } else {
        double* bl = V + 2 * (size_t)l;
        bl[0] = b0; bl[1] = b1;
        double* hl = V + 2 * (size_t)L + 4 * (size_t)l;
        hl[0] = h00; hl[1] = h01; hl[2] = h01; hl[3] = h11;
      }
    }
  }
  // final chi2 of this replica: fixed-order sum of the pose kernel's block partials
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    double t = 0;
    for (int k = threadIdx.x; k < chi2_nblocks; k += 32) t += a.chi2_part[(size_t)r * a.chi2_blocks + k];
    for (int o = 16; o; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    // slot k holds chi2 at the state reached after k iterations (slot 0 = initial estimate)
    const int slot = a.status[2 * r + 1];
    if (threadIdx.x == 0 && slot < a.chi2_cap) a.chi2[(size_t)r * a.chi2_cap + slot] = t;
  }
}

// Second half of the peer exchange: wait until every rank has published `epoch`, then
// V[landmark part] = sum over ranks (ascending) of their partial blocks.  A rank that never arrives
// (a peer died) does not hang the GPU: after `timeout_ns` the kernel sets *err (sticky, read by
// slam_b200_xchg_error), raises the fail flag of the graph (status[0]: factorise/update then leave the
// state alone and graph_finish / graph_optimize report 0 iterations, the g2o failure convention) and
// writes ZEROS into the landmark part of V instead of stale or partial blocks.
__global__ void __launch_bounds__(256)
lm_exchange_sum_kernel(char* const* __restrict__ peer_tab, int world, int rank, int cap, int parity,
                       unsigned long long epoch, const int* __restrict__ ranges, int L, double* __restrict__ V,
                       int* err, int* status, unsigned long long timeout_ns) {
  __shared__ int go;
  // This kernel follows the landmark kernel on the stream, so the rank's slot is complete: CTA 0
  // publishes `epoch` in every rank's flag word (the only remote stores of the exchange).  The grid
  // may exceed what is resident at once (L above ~300k landmarks): correctness does not need full
  // residency, only that CTA 0 gets an SM -- thread blocks are dispatched in index order, so CTA 0 is
  // among the first wave and never waits behind the spinning ones; every spinning CTA only waits for
  // stores made by the CTA 0s of the OTHER ranks' grids.
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    reinterpret_cast<volatile unsigned long long*>(peer_tab[threadIdx.x])[rank] = epoch;
  }
  if (threadIdx.x == 0) {
    const volatile unsigned long long* flags = reinterpret_cast<const volatile unsigned long long*>(peer_tab[rank]);
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int ok = 1;
    for (int s = 0; s < world && ok; s++) {
      while (flags[s] < epoch) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) { ok = 0; break; }
      }
    }
    if (!ok) { *err = 1; status[0] = 1; }
    __threadfence();
    go = ok;
  }
  __syncthreads();
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  double b0 = 0, b1 = 0, h0 = 0, h1 = 0, h2 = 0, h3 = 0;
  for (int s = 0; go && s < world; s++) {
    const int l0 = ranges[2 * s], l1 = ranges[2 * s + 1];
    if (l < l0 || l >= l1) continue;
    const double* slot = xchg_slot(peer_tab[s], cap, parity);  // remote for s != rank: read over NVLink
    const int k = l - l0;
    const double2 b = __ldcg(reinterpret_cast<const double2*>(slot + 2 * (size_t)k));
    const double2* hs = reinterpret_cast<const double2*>(slot + 2 * (size_t)cap + 4 * (size_t)k);
    const double2 ha = __ldcg(hs), hb = __ldcg(hs + 1);
    b0 += b.x; b1 += b.y; h0 += ha.x; h1 += ha.y; h2 += hb.x; h3 += hb.y;
  }
  *reinterpret_cast<double2*>(V + 2 * (size_t)l) = make_double2(b0, b1);
  double2* hv = reinterpret_cast<double2*>(V + 2 * (size_t)L + 4 * (size_t)l);
  hv[0] = make_double2(h0, h1);
  hv[1] = make_double2(h2, h3);
}

// SparseOptimizer::update: VertexSE2::oplusImpl (additive, normalised angle) / VertexPointXY
__global__ void update_kernel(int P, int L, long estStride, double* est_all, const double* x_all, int n,
                              const int* __restrict__ pose_boff, const int* __restrict__ lm_boff,
                              int* status) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (status[2 * r] != 0) return;  // factorisation failed: leave the state alone
  double* est = est_all + (size_t)r * estStride;
  const double* x = x_all + (size_t)r * n;
  if (v < P) {
    int o = pose_boff[v];
    if (o >= 0) {
      est[v] += x[o];
      est[P + v] += x[o + 1];
      est[2 * P + v] = normalize_theta_dev(est[2 * P + v] + x[o + 2]);
    }
  } else if (v < P + L) {
    int l = v - P;
    int o = lm_boff[l];
    if (o >= 0) {
      est[3 * P + l] += x[o];
      est[3 * P + L + l] += x[o + 1];
    }
  }
  if (v == 0) status[2 * r + 1] += 1;
}

// Structure arrays go up through ONE pinned staging buffer: ~40 small copies from pageable vectors
// would each be staged synchronously by the driver (measured 16 ms for the 10-lap graph).
// The list is filled and flushed inside one graph_build_structure call; thread_local so that contexts
// prepared from different host threads do not share it.
struct PendingUpload { void* dst; const void* src; size_t bytes; };
thread_local std::vector<PendingUpload> g_uploads;

template <class T>
int upload_vec(slam_b200_ctx* c, DevBuf<T>& d, const std::vector<T>& h) {
  SLAM_CUDA_TRY(c, d.exact(h.size()));
  if (!h.empty()) g_uploads.push_back({d.p, h.data(), sizeof(T) * h.size()});
  return 0;
}

int flush_uploads(slam_b200_ctx* c) {
  size_t total = 0;
  for (auto& u : g_uploads) total += (u.bytes + 255) & ~(size_t)255;
  SLAM_CUDA_TRY(c, c->pin_stage.reserve(total + 256));
  size_t off = 0;
  for (auto& u : g_uploads) {
    std::memcpy(c->pin_stage.p + off, u.src, u.bytes);
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(u.dst, c->pin_stage.p + off, u.bytes, cudaMemcpyHostToDevice, c->stream));
    off += (u.bytes + 255) & ~(size_t)255;
  }
  g_uploads.clear();
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
}

}  // namespace

int graph_enqueue_update(slam_b200_ctx* c) {
  DeviceSystem& D = *c->sys;
  int nv = D.P + D.L;
  dim3 grid((nv + 255) / 256, D.R);
  update_kernel<<<grid, 256, 0, c->stream>>>(D.P, D.L, D.estStride, D.est.p, D.x.p, D.n, D.pose_boff.p,
                                            D.lm_boff.p, D.status.p);
  c->launches++;
  SLAM_CUDA_TRY(c, cudaGetLastError());
  return 0;
}

void graph_shard_landmarks(DeviceSystem& D, int p0, int p1, int* l0, int* l1) {
  if (D.shard_p0 != p0 || D.shard_p1 != p1) {
    int lo = D.L, hi = -1;
    const int e0 = D.el_start_host[p0], e1 = D.el_start_host[p1];
    for (int e = e0; e < e1; e++) { lo = std::min(lo, D.s_lm_host[e]); hi = std::max(hi, D.s_lm_host[e]); }
    D.shard_p0 = p0; D.shard_p1 = p1;
    D.shard_l0 = hi < 0 ? 0 : lo;
    D.shard_l1 = hi < 0 ? 0 : hi + 1;
  }
  *l0 = D.shard_l0;
  *l1 = D.shard_l1;
}

int graph_enqueue_assemble(slam_b200_ctx* c, int p0, int p1, bool chi2_only, bool peer) {
  DeviceSystem& D = *c->sys;
  AsmArgs a;
  a.P = D.P; a.L = D.L; a.Eo = D.Eo; a.El = D.El;
  a.estStride = D.estStride; a.measStride = D.measStride; a.nV = D.nV;
  a.est = D.est.p; a.meas = D.meas.p; a.V = D.V.p;
  a.pose_free = D.pose_free.p; a.lm_free = D.lm_free.p;
  a.el_start = D.el_start.p; a.el_info = D.el_info.p; a.el_rec = D.el_rec.p;
  a.lm_start = D.lm_start.p; a.lm_edges = D.lm_edges.p; a.lmo_pose = D.lmo_pose.p; a.lmo_info = D.lmo_info.p;
  a.trig = D.trig.p;
  a.eo_i = D.eo_i.p; a.eo_j = D.eo_j.p; a.eo_slot = D.eo_slot.p; a.eo_flags = D.eo_flags.p;
  a.po_start = D.po_start.p; a.po_list = D.po_list.p; a.eo_info = D.eo_info.p;
  a.chi2_part = D.chi2_part.p; a.chi2_blocks = D.chi2_blocks;
  a.chi2 = D.chi2.p; a.chi2_cap = D.chi2_cap; a.status = D.status.p;
  int np = std::max(0, p1 - p0);
  int nblk = (np + ASM_THREADS - 1) / ASM_THREADS;
  if (nblk > 0) {
    dim3 grid(nblk, D.R);
    if (chi2_only) assemble_pose_kernel<true, false><<<grid, ASM_THREADS, 0, c->stream>>>(a, p0, p1);
    else {
      // SLAM_B200_ASM_VARIANT = 10 * depth + minBlocksPerSM selects a pipelined variant (default 24);
      // 0 = the original kernel (kept for A/B measurements: profiles/tools/asm_ab.py).  Measured on
      // B200 (profiles/r01_asm_ab.log): config 5 0.562 -> 0.484 ms, config 3 x 2,048 0.943 -> 0.882 ms;
      // the variants capped at 80 registers (minBlocksPerSM 6) spill and are slower than the original.
      const char* ev = getenv("SLAM_B200_ASM_VARIANT");  // read per call: the A/B tool switches it inside one process
      const int variant = ev ? atoi(ev) : ASM_DEFAULT_VARIANT;
      switch (variant) {
        case 14: assemble_pose_pipe_kernel<1, 4><<<grid, ASM_THREADS, 0, c->stream>>>(a, p0, p1); break;
        case 24: assemble_pose_pipe_kernel<2, 4><<<grid, ASM_THREADS, 0, c->stream>>>(a, p0, p1); break;
        default: assemble_pose_kernel<false, true><<<grid, ASM_THREADS, 0, c->stream>>>(a, p0, p1); break;
      }
    }
    c->launches++;
  }
  // landmarks touched by this pose range (a shard of a large graph sees a small, contiguous-ish part
  // of the landmarks); the untouched part of the landmark blocks must read as zero for the reduction
  int l_first = 0, l_end = D.L;
  PeerArgs px = {};
  if (peer) {
    PeerExchange& X = D.xchg;
    if (chi2_only || D.R != 1 || !X.connected) return SLAM_B200_E_STATE;
    graph_shard_landmarks(D, p0, p1, &l_first, &l_end);
    if (l_first != X.ranges_host[2 * X.rank] || l_end != X.ranges_host[2 * X.rank + 1] || l_end - l_first > X.cap)
      return SLAM_B200_E_ARG;  // the pose range differs from the one the exchange was connected for
    X.epoch++;
    px.peer_tab = X.peer_tab.p; px.world = X.world; px.rank = X.rank; px.cap = X.cap;
    px.parity = (int)(X.epoch & 1); px.epoch = X.epoch;
  } else if (!chi2_only && (p0 > 0 || p1 < D.P) && D.R == 1) {
    graph_shard_landmarks(D, p0, p1, &l_first, &l_end);
    SLAM_CUDA_TRY(c, cudaMemsetAsync(D.V.p, 0, sizeof(double) * 6 * (size_t)D.L, c->stream));
  }
  // lanes per landmark from the mean number of observers (SLAM_B200_LM_GROUP overrides: 8, 16, 32)
  int G = 32;
  {
    const double mean_obs = (double)D.El / std::max(D.L, 1);
    // measured (profiles/r01_asm_ab.log): 36 observers (1M-pose graph) 0.489 / 0.437 / 0.417 ms and 24
    // observers (1-lap replicas) 0.888 / 0.769 / 0.722 ms for G = 32 / 16 / 8 (4 lanes: no further gain);
    // 279 observers (10-lap graph) 36 / 42 / 54 us
    if (mean_obs <= 48.0) G = 8;
    else if (mean_obs <= 128.0) G = 16;
    if (const char* ev = getenv("SLAM_B200_LM_GROUP")) { int v = atoi(ev); if (v == 8 || v == 16 || v == 32) G = v; }
  }
  const int lm_per_block = LM_PER_BLOCK * (32 / G);
  dim3 gl(std::max(1, (l_end - l_first + lm_per_block - 1) / lm_per_block), D.R);
  // a landmark with more than one G-edge step of observers (long tracks, many laps) is worth the
  // two-steps-in-flight loop; with fewer the plain loop is faster (measured)
  const bool unroll2 = D.El > (long)G * std::max(D.L, 1);
#define LM_LAUNCH(C, U, PE, GG) assemble_landmark_kernel<C, U, PE, GG><<<gl, ASM_THREADS, 0, c->stream>>>(a, p0, p1, nblk, l_first, l_end, px)
#define LM_LAUNCH_G(C, U, PE) do { if (G == 8) LM_LAUNCH(C, U, PE, 8); else if (G == 16) LM_LAUNCH(C, U, PE, 16); else LM_LAUNCH(C, U, PE, 32); } while (0)
  if (chi2_only) LM_LAUNCH(true, false, false, 32);
  else if (peer && unroll2) LM_LAUNCH_G(false, true, true);
  else if (peer) LM_LAUNCH_G(false, false, true);
  else if (unroll2) LM_LAUNCH_G(false, true, false);
  else LM_LAUNCH_G(false, false, false);
#undef LM_LAUNCH_G
#undef LM_LAUNCH
  c->launches++;
  if (peer) {
    PeerExchange& X = D.xchg;
    lm_exchange_sum_kernel<<<(D.L + 255) / 256, 256, 0, c->stream>>>(X.peer_tab.p, X.world, X.rank, X.cap, px.parity, X.epoch,
                                                                   X.ranges.p, D.L, D.V.p, X.err.p, D.status.p, X.timeout_ns);
    c->launches++;
  }
  SLAM_CUDA_TRY(c, cudaGetLastError());
  if (!chi2_only) D.assembled = true;
  return 0;
}

// ================================================================================================
// host structure pass
// ================================================================================================
int graph_build_structure(slam_b200_ctx* c) {
  HostGraph& g = c->g;
  if (!c->sys) c->sys = new DeviceSystem();
  DeviceSystem& D = *c->sys;
  if (D.structure_version == g.structure_version && D.assembly_only == c->assembly_only &&
      D.batch_ordering == c->batch_ordering)
    return D.n;
  D.assembly_only = c->assembly_only;
  D.batch_ordering = c->batch_ordering;
  auto t0 = std::chrono::steady_clock::now();
  const int P = g.P(), L = g.L(), Eo = g.Eo(), El = g.El();
  D.P = P; D.L = L; D.Eo = Eo; D.El = El;
  // profiler ranges of the three host phases (they share locals, hence push/pop instead of scopes); the
  // guard pops whatever is open on any return
  struct PhaseRanges {
    int open = 0;
    void next(const char* name) { if (open) nvtxRangePop(); nvtxRangePushA(name); open = 1; }
    ~PhaseRanges() { if (open) nvtxRangePop(); }
  } phase;
  phase.next("slam_b200/host structure pass");
  static const bool dbg_on = getenv("SLAM_B200_SYM_DEBUG") != nullptr;
  auto tdb = t0;
  auto dbg = [&](const char* what) {
    if (!dbg_on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[structure] %-28s %.4f s\n", what, std::chrono::duration<double>(now - tdb).count());
    tdb = now;
  };
  // ---- active set (initializeOptimization): edges whose vertices are not all fixed ----
  std::vector<char> pose_act(P, 0), lm_act(L, 0);
  std::vector<int> el_flags(El, 0), eo_flags(Eo, 0);
  for (int e = 0; e < El; e++)
    if (!(g.pose_fixed[g.el_p[e]] && g.lm_fixed[g.el_l[e]])) {
      el_flags[e] = EF_ACTIVE;
      pose_act[g.el_p[e]] = 1;
      lm_act[g.el_l[e]] = 1;
    }
  for (int e = 0; e < Eo; e++)
    if (!(g.pose_fixed[g.eo_i[e]] && g.pose_fixed[g.eo_j[e]])) {
      eo_flags[e] = EF_ACTIVE;
      pose_act[g.eo_i[e]] = pose_act[g.eo_j[e]] = 1;
    }
  // ---- Hessian index mapping (buildIndexMapping): non-fixed active vertices by ascending id ----
  std::vector<std::pair<int, int>> byId;  // (id, (local<<1)|is_lm)
  byId.reserve((size_t)P + L);
  // landmarks first: in the reference's numbering (cone ids from 0, pose ids from 1000 up) that is already the
  // ascending order, and the sort below is skipped
  for (int l = 0; l < L; l++)
    if (lm_act[l] && !g.lm_fixed[l]) byId.push_back({g.lm_id[l], (l << 1) | 1});
  for (int p = 0; p < P; p++)
    if (pose_act[p] && !g.pose_fixed[p]) byId.push_back({g.pose_id[p], p << 1});
  if (!std::is_sorted(byId.begin(), byId.end())) std::sort(byId.begin(), byId.end());
  const int nb = (int)byId.size();
  D.nb = nb;
  D.pose_b.assign(P, -1);
  D.lm_b.assign(L, -1);
  D.blk_hidx.assign(nb, 0);
  D.blk_kind_local.assign(nb, 0);
  std::vector<int> dim(nb);
  {
    int off = 0;
    for (int b = 0; b < nb; b++) {
      int kl = byId[b].second;
      D.blk_kind_local[b] = kl;
      if (kl & 1) { D.lm_b[kl >> 1] = b; dim[b] = 2; }
      else { D.pose_b[kl >> 1] = b; dim[b] = 3; }
      D.blk_hidx[b] = off;
      off += dim[b];
    }
    D.n = off;
  }
  dbg("active set + index map");
  // ---- V layout ----
  const long base_off = 6L * L + 12L * P;
  D.hoff_diag.assign(nb, 0);
  for (int b = 0; b < nb; b++) {
    int kl = D.blk_kind_local[b];
    D.hoff_diag[b] = (kl & 1) ? (int)(2L * L + 4L * (kl >> 1)) : (int)(6L * L + 3L * P + 9L * (kl >> 1));
  }
  // ---- block structure (buildStructure): one slot per distinct free vertex pair ----
  // Only what fixes the PATTERN is computed before the symbolic analysis is started (it runs on a helper
  // thread from then on); payload arrays, incidence lists and the first part of the upload follow while it runs.
  D.off_a.clear(); D.off_b.clear(); D.hoff_off.clear();
  long cursor = base_off;
  auto new_slot = [&](int ba, int bb) -> int {
    int lo = std::min(ba, bb), hi = std::max(ba, bb);
    int k = (int)D.off_a.size();
    D.off_a.push_back(lo);
    D.off_b.push_back(hi);
    D.hoff_off.push_back((int)cursor);
    cursor += dim[lo] * dim[hi];
    return k;
  };
  // landmark edges sorted by pose (stable): the order performSLAM creates them in
  D.el_perm.resize(El);
  std::vector<int> el_start(P + 1, 0);
  for (int e = 0; e < El; e++) el_start[g.el_p[e] + 1]++;
  for (int p = 0; p < P; p++) el_start[p + 1] += el_start[p];
  {
    std::vector<int> cur(el_start.begin(), el_start.end() - 1);
    for (int e = 0; e < El; e++) D.el_perm[cur[g.el_p[e]]++] = e;
  }
  std::vector<int> s_pose(El), s_lm(El), s_slot(El, -1), s_flags(El, 0);
  {
    // A (pose, landmark) pair can only repeat inside one pose's run of edges, so duplicates are found
    // by a linear look-back over that run (a handful of edges) -- no global hash for the 10^5..10^7
    // pose-landmark blocks.
    D.off_a.reserve((size_t)El + Eo);
    D.off_b.reserve((size_t)El + Eo);
    D.hoff_off.reserve((size_t)El + Eo);
    std::vector<int> run_slot(El, -1);  // slot index of every sorted edge (or -1)
    int runStart = 0;
    for (int q = 0; q < El; q++) {
      int e = D.el_perm[q];
      int p = g.el_p[e], l = g.el_l[e];
      if (q > 0 && p != s_pose[q - 1]) runStart = q;
      s_pose[q] = p;
      s_lm[q] = l;
      int fl = el_flags[e];
      if ((fl & EF_ACTIVE) && D.pose_b[p] >= 0 && D.lm_b[l] >= 0) {
        int k = -1;
        for (int t = runStart; t < q; t++)
          if (s_lm[t] == l && run_slot[t] >= 0) { k = run_slot[t]; break; }
        if (k < 0) {
          k = new_slot(D.pose_b[p], D.lm_b[l]);
          fl |= EF_FIRST;
        }
        run_slot[q] = k;
        s_slot[q] = D.hoff_off[k];
        fl |= EF_OFFDIAG;
        if (D.lm_b[l] < D.pose_b[p]) fl |= EF_TRANS;
      }
      s_flags[q] = fl;
    }
  }
  dbg("landmark edges: sort + slots");
  // pose-pose edges: incidence lists, slots, owner = min(i, j).  Edges of one pair are all processed by thread
  // min(i, j) in incidence (= edge) order, so "has this pair been seen" is a look-back over the owner's earlier
  // incidence entries (two or three for a pose chain) -- no hash map; a hub pose falls back to one.
  std::vector<int> po_start(P + 1, 0), po_list(2 * (size_t)Eo), eo_slot(Eo, -1);
  for (int e = 0; e < Eo; e++) { po_start[g.eo_i[e] + 1]++; po_start[g.eo_j[e] + 1]++; }
  for (int p = 0; p < P; p++) po_start[p + 1] += po_start[p];
  {
    std::vector<int> cur(po_start.begin(), po_start.end() - 1);
    for (int e = 0; e < Eo; e++) {
      po_list[cur[g.eo_i[e]]++] = (e << 1);
      po_list[cur[g.eo_j[e]]++] = (e << 1) | 1;
    }
    std::vector<int> eo_k(Eo, -1);  // slot index of every edge that has one
    std::unordered_map<uint64_t, int> hub;  // only for owners with long incidence lists
    for (int e = 0; e < Eo; e++) {
      int i = g.eo_i[e], j = g.eo_j[e];
      if ((eo_flags[e] & EF_ACTIVE) && D.pose_b[i] >= 0 && D.pose_b[j] >= 0) {
        const int owner = std::min(i, j), partner = std::max(i, j);
        int k = -1;
        const int deg = po_start[owner + 1] - po_start[owner];
        if (deg <= 32) {
          for (int t = po_start[owner]; t < po_start[owner + 1]; t++) {
            const int e2 = po_list[t] >> 1;
            if (e2 >= e) break;  // incidence entries are in edge order
            if (eo_k[e2] >= 0 && std::min(g.eo_i[e2], g.eo_j[e2]) == owner && std::max(g.eo_i[e2], g.eo_j[e2]) == partner) { k = eo_k[e2]; break; }
          }
        } else {
          const uint64_t key = ((uint64_t)(uint32_t)owner << 32) | (uint32_t)partner;
          auto it = hub.find(key);
          if (it != hub.end()) k = it->second;
        }
        if (k < 0) {
          k = new_slot(D.pose_b[i], D.pose_b[j]);
          eo_flags[e] |= EF_FIRST;
          if (deg > 32) hub.emplace(((uint64_t)(uint32_t)owner << 32) | (uint32_t)partner, k);
        }
        eo_k[e] = k;
        eo_slot[e] = D.hoff_off[k];
        eo_flags[e] |= EF_OFFDIAG;
        if (D.pose_b[j] < D.pose_b[i]) eo_flags[e] |= EF_TRANS;
      }
    }
  }
  dbg("pose-pose edges");
  D.nV = cursor;
  D.t_structure = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // ---- symbolic analysis (helper thread) ----
  // Nested-dissection region size: about a tenth of the graph (measured optimum on the 1-lap and
  // 10-lap trackdrive graphs: short assembly tree, still enough regions to order in parallel), but
  // never so small that a hub vertex (a landmark seen from hundreds of poses) dominates a region --
  // below ~2.5x the largest degree the separators degenerate and fill explodes.
  int leaf;
  {
    std::vector<int> deg(nb, 0);
    for (size_t k = 0; k < D.off_a.size(); k++) { deg[D.off_a[k]]++; deg[D.off_b[k]]++; }
    int maxdeg = 0;
    for (int b = 0; b < nb; b++) maxdeg = std::max(maxdeg, deg[b]);
    leaf = std::min(std::max(std::max(nb / 10, 3 * maxdeg), 64), 2048);
    // a batch of replicas is throughput-bound, not latency-bound: prefer the ordering with the least
    // fill (larger regions, more of the graph ordered by minimum degree)
    if (c->batch_ordering) leaf = std::min(std::max(1024, 3 * maxdeg), 2048);
  }
  if (const char* s = getenv("SLAM_B200_ND_LEAF")) leaf = std::max(1, atoi(s));
  std::thread sym_thread;
  std::exception_ptr sym_error;
  // the analysis signals when the tree and the fronts are final; its last stage (the assembly entries) then runs
  // beside the launch lists built below from them
  std::mutex sym_m;
  std::condition_variable sym_cv;
  bool sym_ready = false, sym_structure_ok = false;
  const std::function<void(bool)> sym_structure_ready = [&](bool ok) {
    { std::lock_guard<std::mutex> lk(sym_m); sym_ready = true; sym_structure_ok = ok; }
    sym_cv.notify_all();
  };
  struct JoinGuard {  // an early return below must not leave the helper running on locals of this frame
    std::thread& t;
    ~JoinGuard() { if (t.joinable()) t.join(); }
  } join_guard{sym_thread};
  static const bool no_overlap = getenv("SLAM_B200_NO_HOST_OVERLAP") != nullptr;
  if (c->assembly_only) {
    // linearise + assemble only (config 5 measures the edge-partitioned assembly; the solve is
    // reported separately): no ordering, no fronts
    D.sym = Symbolic();
    D.sym.nb = nb;
    D.sym.n = D.n;
    D.sym.boff = D.blk_hidx;  // solver order = g2o order
  } else {
    auto run_symbolic = [&]() {
      try {
        nvtxRangePushA("slam_b200/symbolic analysis");
        symbolic_analyze(nb, dim.data(), (int)D.off_a.size(), D.off_a.data(), D.off_b.data(), D.hoff_diag.data(),
                         D.hoff_off.data(), leaf, D.sym, 0, &sym_structure_ready);
        nvtxRangePop();
      } catch (...) {
        sym_error = std::current_exception();
      }
    };
    if (no_overlap) run_symbolic();
    else sym_thread = std::thread(run_symbolic);
  }
  // ---- while the analysis runs: payload arrays of the assembly kernels and their upload ----
  phase.next("slam_b200/assembly structure + upload (overlaps the symbolic analysis)");
  auto tov0 = std::chrono::steady_clock::now();
  std::vector<double> s_info(3 * (size_t)El);
  for (int q = 0; q < El; q++) {
    const int e = D.el_perm[q];
    s_info[q] = g.el_info[3 * (size_t)e];
    s_info[El + (size_t)q] = g.el_info[3 * (size_t)e + 1];
    s_info[2 * (size_t)El + q] = g.el_info[3 * (size_t)e + 2];
  }
  // CSR landmark -> sorted edge positions
  std::vector<int> lm_start(L + 1, 0), lm_edges(El);
  for (int q = 0; q < El; q++) lm_start[s_lm[q] + 1]++;
  for (int l = 0; l < L; l++) lm_start[l + 1] += lm_start[l];
  {
    std::vector<int> cur(lm_start.begin(), lm_start.end() - 1);
    for (int q = 0; q < El; q++) lm_edges[cur[s_lm[q]]++] = q;
  }
  // landmark-ordered payload of the landmark kernel: pose of every edge (-1 = inactive) and its
  // information matrix, in the order of lm_edges, so the lanes of a warp read consecutive memory
  std::vector<int> lmo_pose(El);
  std::vector<double> lmo_info(3 * (size_t)El);
  for (int q = 0; q < El; q++) {
    const int e = lm_edges[q];  // pose-sorted position
    lmo_pose[q] = (s_flags[e] & EF_ACTIVE) ? s_pose[e] : -1;
    lmo_info[q] = s_info[e];
    lmo_info[El + (size_t)q] = s_info[El + (size_t)e];
    lmo_info[2 * (size_t)El + q] = s_info[2 * (size_t)El + e];
  }
  D.lm_order = lm_edges;
  D.s_lm_host = s_lm;
  D.el_start_host = el_start;
  D.shard_p0 = D.shard_p1 = -1;
  std::vector<double> eo_info(6 * (size_t)Eo);
  for (int e = 0; e < Eo; e++)
    for (int k = 0; k < 6; k++) eo_info[(size_t)k * Eo + e] = g.eo_info[6 * (size_t)e + k];
  std::vector<unsigned char> pose_free(P, 0), lm_free(L, 0);
  for (int b = 0; b < nb; b++) {
    const int kl = D.blk_kind_local[b];
    if (kl & 1) lm_free[kl >> 1] = 1;
    else pose_free[kl >> 1] = 1;
  }
  std::vector<int4> el_rec_h(El);
  for (int q = 0; q < El; q++) el_rec_h[q] = make_int4(s_pose[q], s_lm[q], s_slot[q], s_flags[q]);
  dbg("assembly payload (overlapped)");
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  g_uploads.clear();
  {
    int rc = 0;
    rc |= upload_vec(c, D.pose_free, pose_free);
    rc |= upload_vec(c, D.lm_free, lm_free);
    rc |= upload_vec(c, D.el_start, el_start);
    rc |= upload_vec(c, D.el_rec, el_rec_h);
    rc |= upload_vec(c, D.el_info, s_info);
    rc |= upload_vec(c, D.lm_start, lm_start);
    rc |= upload_vec(c, D.lm_edges, lm_edges);
    rc |= upload_vec(c, D.lmo_pose, lmo_pose);
    rc |= upload_vec(c, D.lmo_info, lmo_info);
    rc |= upload_vec(c, D.eo_i, g.eo_i);
    rc |= upload_vec(c, D.eo_j, g.eo_j);
    rc |= upload_vec(c, D.eo_slot, eo_slot);
    rc |= upload_vec(c, D.eo_flags, eo_flags);
    rc |= upload_vec(c, D.po_start, po_start);
    rc |= upload_vec(c, D.po_list, po_list);
    rc |= upload_vec(c, D.eo_info, eo_info);
    if (rc) { g_uploads.clear(); return SLAM_B200_E_CUDA; }
    if (int frc = flush_uploads(c)) return frc;
  }
  dbg("assembly upload (overlapped)");
  const double t_overlapped = std::chrono::duration<double>(std::chrono::steady_clock::now() - tov0).count();
  if (sym_thread.joinable()) {
    std::unique_lock<std::mutex> lk(sym_m);
    sym_cv.wait(lk, [&] { return sym_ready; });  // fired on success and on the way out of an exception alike
  }
  auto sym_join = [&]() {  // the analysis has finished altogether (assembly entries, timing); rethrows its exception
    if (sym_thread.joinable()) sym_thread.join();
    if (sym_error) std::rethrow_exception(sym_error);
  };
  if (!sym_structure_ok) sym_join();  // no helper thread (inline / assembly only) or the analysis failed early
  dbg("wait for the symbolic analysis");
  phase.next("slam_b200/launch lists");
  Symbolic& S = D.sym;
  if (S.lptr.empty()) { S.lptr.assign(1, 0); S.uptr.assign(1, 0); S.rows_ptr.assign(1, 0); S.child_ptr.assign(1, 0); S.asm_ptr.assign(1, 0); }
  D.nL = S.lptr[S.nf];
  D.nU = S.uptr[S.nf];
  D.nUvec = S.rows_ptr[S.nf];
  // per-level launch lists: fronts whose dense frontal matrix fits in shared memory vs the rest
  // room behind the front for the row map and the scaled panel (solver.cu: factor_extra_smem)
  const size_t smem_limit = (size_t)std::max(0, c->max_smem_optin - 16384);
  std::vector<int> launch_list;
  std::vector<long> fbig(S.nf, -1);
  D.levels.assign(S.nlevels, LevelLaunch());
  D.nFbig = 0;
  for (int lv = 0; lv < S.nlevels; lv++) {
    LevelLaunch& LL = D.levels[lv];
    LL.list_off = (int)launch_list.size();
    std::vector<int> small, big;
    for (int f = S.level_ptr[lv]; f < S.level_ptr[lv + 1]; f++) {
      size_t fs = (size_t)S.npiv[f] + S.nupd[f];
      size_t need = fs * (size_t)front_ld((int)fs) * sizeof(double);  // + the right-hand-side row, padded stride (factor2_kernel)
      LL.max_fs = std::max(LL.max_fs, (int)fs);
      if (fs <= 64 && need <= smem_limit) {
        launch_list.push_back(f);
        LL.n_tiny++;
        LL.smem_tiny = std::max(LL.smem_tiny, need);
        LL.max_fs_tiny = std::max(LL.max_fs_tiny, (int)fs);
      } else if (need <= smem_limit) {
        small.push_back(f);
        LL.smem_factor = std::max(LL.smem_factor, need);
      } else {
        big.push_back(f);
        fbig[f] = D.nFbig;
        D.nFbig += (long)(fs * (fs + 1));
      }
      size_t sneed = (fs * S.npiv[f] + fs) * sizeof(double);
      LL.smem_solve = std::max(LL.smem_solve, std::min(sneed, smem_limit));
    }
    {
      auto fsz = [&](int f) { return S.npiv[f] + S.nupd[f]; };
      std::stable_sort(launch_list.begin() + LL.list_off, launch_list.end(), [&](int a, int b) { return fsz(a) < fsz(b); });
      for (size_t q = LL.list_off; q < launch_list.size(); q++) {
        const int fs = fsz(launch_list[q]);
        const int cls = fs <= 40 ? 0 : fs <= 48 ? 1 : fs <= 56 ? 2 : 3;
        LL.tiny_cls_n[cls]++;
        LL.tiny_cls_fs[cls] = std::max(LL.tiny_cls_fs[cls], fs);
      }
    }
    for (int f : small) launch_list.push_back(f);
    for (int f : big) launch_list.push_back(f);
    LL.n_small = (int)small.size();
    LL.n_big = (int)big.size();
  }
  D.launch_list_host = launch_list;
  // forward-solve gather lists (by destination row of the parent front, children in order)
  std::vector<int> frow_ptr(S.nf + 1, 0), gather_ptr, gather_src;
  for (int f = 0; f < S.nf; f++) frow_ptr[f + 1] = frow_ptr[f] + S.npiv[f] + S.nupd[f] + 1;
  gather_ptr.assign(frow_ptr[S.nf], 0);
  {
    // counting sort per front (children in order, rows ascending inside a child): every update row of
    // a non-root front lands in exactly one row of its parent
    gather_src.assign((size_t)S.rows_ptr[S.nf], 0);
    std::vector<int> cur;
    int base = 0;
    for (int f = 0; f < S.nf; f++) {
      const int fs = S.npiv[f] + S.nupd[f];
      int* gp = gather_ptr.data() + frow_ptr[f];  // fs + 1 entries
      std::fill(gp, gp + fs + 1, 0);
      for (int ci = S.child_ptr[f]; ci < S.child_ptr[f + 1]; ci++) {
        const int ch = S.children[ci];
        for (int q = S.rows_ptr[ch]; q < S.rows_ptr[ch + 1]; q++) gp[S.rel[q] + 1]++;
      }
      gp[0] = base;
      for (int i = 0; i < fs; i++) gp[i + 1] += gp[i];
      cur.assign(gp, gp + fs);
      for (int ci = S.child_ptr[f]; ci < S.child_ptr[f + 1]; ci++) {
        const int ch = S.children[ci];
        for (int q = S.rows_ptr[ch]; q < S.rows_ptr[ch + 1]; q++) gather_src[cur[S.rel[q]]++] = q;
      }
      base = gp[fs];
    }
    gather_src.resize((size_t)base);
  }
  // solver scalar -> V offset of the rhs entry; vertex -> solver offset
  std::vector<int> solver2v(D.n), pose_boff(P, -1), lm_boff(L, -1);
  for (int b = 0; b < nb; b++) {
    int kl = D.blk_kind_local[b];
    int so = S.boff[b];
    if (kl & 1) {
      int l = kl >> 1;
      lm_boff[l] = so;
      solver2v[so] = 2 * l;
      solver2v[so + 1] = 2 * l + 1;
    } else {
      int p = kl >> 1;
      pose_boff[p] = so;
      for (int k = 0; k < 3; k++) solver2v[so + k] = 6 * L + 3 * p + k;
    }
  }
  sym_join();  // the assembly entries (tile plan, upload) and the analysis' own clock are needed from here on
  // ---- tiled path for replica batches (tileplan.h) ----
  D.tile_path = false;
  D.tile = TilePlan();
  D.tile_launches.clear();
  std::vector<int> tile_list;
  std::vector<int2> tile_items_h;
  if (c->batch_ordering && !c->assembly_only && !getenv("SLAM_B200_NO_TILE_PATH")) {
    tile_plan_build(S, solver2v, D.tile);
    if (D.tile.ok && D.nV <= (long)INT32_MAX) {
      D.tile_path = true;
      D.nL = D.tile.nF();   // Lv = tile storage of every front (L tiles + Schur-complement tiles)
      D.nU = 0;
      D.nUvec = 0;
      // classes of tile rows: a launch sizes its shared memory for the largest front of its class
      // fronts of <= 8 tile rows: one class per T (the register-resident kernel is instantiated per T)
      auto cls_of = [](int T) { return T <= 8 ? T : T <= 10 ? 9 : 10; };
      for (int lv = 0; lv < S.nlevels; lv++) {
        std::vector<int> fr;
        for (int f = S.level_ptr[lv]; f < S.level_ptr[lv + 1]; f++) fr.push_back(f);
        std::stable_sort(fr.begin(), fr.end(), [&](int a, int b) { return D.tile.T[a] < D.tile.T[b]; });
        size_t q = 0;
        while (q < fr.size()) {
          TileLaunch TL;
          TL.level = lv;
          TL.list_off = (int)tile_list.size();
          const int cls = cls_of(D.tile.T[fr[q]]);
          while (q < fr.size() && cls_of(D.tile.T[fr[q]]) == cls) {
            TL.T = std::max(TL.T, D.tile.T[fr[q]]);
            tile_list.push_back(fr[q]);
            TL.count++;
            q++;
          }
          D.tile_launches.push_back(TL);
        }
      }
      tile_items_h.resize(D.tile.items.size());
      for (size_t k = 0; k < D.tile.items.size(); k++) tile_items_h[k] = make_int2(D.tile.items[k].src, D.tile.items[k].dst);
    }
  }
  D.t_lists = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - D.t_structure -
              std::max(S.seconds, t_overlapped);
  dbg("launch lists");
  // ---- upload the analysis ----
  phase.next("slam_b200/structure upload");
  auto tu0 = std::chrono::steady_clock::now();
  g_uploads.clear();
  int rc = 0;
  rc |= upload_vec(c, D.pose_boff, pose_boff);
  rc |= upload_vec(c, D.lm_boff, lm_boff);
  rc |= upload_vec(c, D.ds.piv0, S.piv0);
  rc |= upload_vec(c, D.ds.npiv, S.npiv);
  rc |= upload_vec(c, D.ds.nupd, S.nupd);
  rc |= upload_vec(c, D.ds.rows_ptr, S.rows_ptr);
  rc |= upload_vec(c, D.ds.upd_rows, S.upd_rows);
  rc |= upload_vec(c, D.ds.rel, S.rel);
  rc |= upload_vec(c, D.ds.child_ptr, S.child_ptr);
  rc |= upload_vec(c, D.ds.children, S.children);
  rc |= upload_vec(c, D.ds.asm_ptr, S.asm_ptr);
  rc |= upload_vec(c, D.ds.asm_entries, S.asm_entries);
  rc |= upload_vec(c, D.ds.solver2v, solver2v);
  rc |= upload_vec(c, D.ds.lptr, S.lptr);
  rc |= upload_vec(c, D.ds.uptr, S.uptr);
  rc |= upload_vec(c, D.ds.fbig, fbig);
  rc |= upload_vec(c, D.ds.launch_list, launch_list);
  rc |= upload_vec(c, D.ds.frow_ptr, frow_ptr);
  rc |= upload_vec(c, D.ds.gather_ptr, gather_ptr);
  rc |= upload_vec(c, D.ds.gather_src, gather_src);
  if (D.tile_path) {
    rc |= upload_vec(c, D.tile_list, tile_list);
    rc |= upload_vec(c, D.tile_item_ptr, D.tile.item_ptr);
    rc |= upload_vec(c, D.tile_item_nv, D.tile.item_nv);
    rc |= upload_vec(c, D.tile_fptr, D.tile.fptr);
    rc |= upload_vec(c, D.tile_items, tile_items_h);
  }
  if (rc) { g_uploads.clear(); return SLAM_B200_E_CUDA; }
  dbg("upload: allocations + records");
  if (int frc = flush_uploads(c)) return frc;
  dbg("upload: staging + copies");
  D.structure_version = g.structure_version;
  D.values_version = 0;
  D.R = 0;  // value arrays must be (re)allocated for the new sizes
  D.drop_graph();
  D.assembled = false;
  D.upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tu0).count();
  D.overlapped_seconds = t_overlapped;
  return D.n;
}

int graph_alloc_values(slam_b200_ctx* c, int R) {
  DeviceSystem& D = *c->sys;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  D.estStride = 3L * D.P + 2L * D.L;
  D.measStride = 4L * D.El + 3L * D.Eo;  // + the landmark-ordered copy of the cone measurements
  D.chi2_cap = 64;
  D.chi2_blocks = std::max(1, (D.P + ASM_THREADS - 1) / ASM_THREADS);
  if (R != D.R) {
    size_t r = (size_t)R;
    SLAM_CUDA_TRY(c, D.est.exact(r * D.estStride));
    SLAM_CUDA_TRY(c, D.meas.exact(r * D.measStride));
    SLAM_CUDA_TRY(c, D.V.exact(r * D.nV));
    SLAM_CUDA_TRY(c, D.Lv.exact(r * D.nL));
    SLAM_CUDA_TRY(c, D.Uv.exact(r * D.nU));
    SLAM_CUDA_TRY(c, D.uvec.exact(r * D.nUvec));
    SLAM_CUDA_TRY(c, D.x.exact(r * D.n));
    SLAM_CUDA_TRY(c, D.Fbig.exact(r * D.nFbig));
    SLAM_CUDA_TRY(c, D.chi2.exact(r * D.chi2_cap));
    SLAM_CUDA_TRY(c, D.chi2_part.exact(r * D.chi2_blocks));
    SLAM_CUDA_TRY(c, D.status.exact(2 * r));
    if (getenv("SLAM_B200_PHASE_CLOCKS")) {
      SLAM_CUDA_TRY(c, D.dbg_clocks.exact(32));
      SLAM_CUDA_TRY(c, cudaMemsetAsync(D.dbg_clocks.p, 0, sizeof(long long) * 32, c->stream));
    }
    if (getenv("SLAM_B200_TIMELINE") && D.sym.nf > 0) {
      SLAM_CUDA_TRY(c, D.timeline.exact(12 * (size_t)D.sym.nf));
      SLAM_CUDA_TRY(c, cudaMemsetAsync(D.timeline.p, 0, sizeof(long long) * 12 * (size_t)D.sym.nf, c->stream));
    }
    SLAM_CUDA_TRY(c, D.est0.exact(r * D.estStride));
    SLAM_CUDA_TRY(c, D.trig.exact(r * 2 * (size_t)D.P));
    D.R = R;
    D.drop_graph();
    D.values_version = 0;
  }
  SLAM_CUDA_TRY(c, cudaMemsetAsync(D.status.p, 0, sizeof(int) * 2 * (size_t)R, c->stream));
  SLAM_CUDA_TRY(c, cudaMemsetAsync(D.chi2.p, 0, sizeof(double) * (size_t)R * D.chi2_cap, c->stream));
  D.iters_enqueued = 0;
  D.assembled = false;
  return 0;
}

// replica 0 <- host graph numbers (estimates + measurements)
int graph_upload_host_values(slam_b200_ctx* c) {
  HostGraph& g = c->g;
  DeviceSystem& D = *c->sys;
  const int P = D.P, L = D.L, Eo = D.Eo, El = D.El;
  size_t ne = (size_t)D.estStride, nm = (size_t)D.measStride;
  SLAM_CUDA_TRY(c, c->pin_d.reserve(ne + nm + 8));
  double* e = c->pin_d.p;
  for (int p = 0; p < P; p++) {
    e[p] = g.pose_est[3 * (size_t)p];
    e[P + p] = g.pose_est[3 * (size_t)p + 1];
    e[2 * (size_t)P + p] = g.pose_est[3 * (size_t)p + 2];
  }
  for (int l = 0; l < L; l++) {
    e[3 * (size_t)P + l] = g.lm_est[2 * (size_t)l];
    e[3 * (size_t)P + L + l] = g.lm_est[2 * (size_t)l + 1];
  }
  double* m = e + ne;
  for (int q = 0; q < El; q++) {
    int o = D.el_perm[q];
    m[q] = g.el_z[2 * (size_t)o];
    m[El + (size_t)q] = g.el_z[2 * (size_t)o + 1];
  }
  for (int k = 0; k < Eo; k++) {
    m[2 * (size_t)El + k] = g.eo_z[3 * (size_t)k];
    m[2 * (size_t)El + Eo + k] = g.eo_z[3 * (size_t)k + 1];
    m[2 * (size_t)El + 2 * (size_t)Eo + k] = g.eo_z[3 * (size_t)k + 2];
  }
  double* ml = m + 2 * (size_t)El + 3 * (size_t)Eo;  // landmark order
  for (int q = 0; q < El; q++) {
    int e = D.lm_order[q];
    ml[q] = m[e];
    ml[El + (size_t)q] = m[El + (size_t)e];
  }
  if (ne) SLAM_CUDA_TRY(c, cudaMemcpyAsync(D.est.p, e, sizeof(double) * ne, cudaMemcpyHostToDevice, c->stream));
  if (nm) SLAM_CUDA_TRY(c, cudaMemcpyAsync(D.meas.p, m, sizeof(double) * nm, cudaMemcpyHostToDevice, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // pinned staging buffer is reused
  D.values_version = g.values_version;
  return 0;
}

void xchg_release(DeviceSystem& D) {
  PeerExchange& X = D.xchg;
  for (int r = 0; r < (int)X.peers.size(); r++)
    if (r != X.rank && X.peers[r]) cudaIpcCloseMemHandle(X.peers[r]);
  if (X.local) cudaFree(X.local);
  X.peer_tab.release(); X.ranges.release(); X.err.release(); X.done.release();
  X = PeerExchange();
}

bool graph_device_estimates(slam_b200_ctx* c, double** est, int* P, int* L) {
  if (!c->sys) return false;
  DeviceSystem& D = *c->sys;
  if (D.R != 1 || !D.est.p || D.structure_version != c->g.structure_version || D.values_version != c->g.values_version ||
      D.P != c->g.P() || D.L != c->g.L())
    return false;
  *est = D.est.p;
  *P = D.P;
  *L = D.L;
  return true;
}

void graph_release(slam_b200_ctx* c) {
  if (!c->sys) return;
  DeviceSystem& D = *c->sys;
  D.pose_free.release(); D.lm_free.release(); D.pose_boff.release(); D.lm_boff.release();
  D.el_start.release(); D.el_rec.release();
  D.el_info.release(); D.lm_start.release(); D.lm_edges.release();
  D.eo_i.release(); D.eo_j.release(); D.eo_slot.release(); D.eo_flags.release(); D.po_start.release();
  D.po_list.release(); D.eo_info.release();
  D.ds.piv0.release(); D.ds.npiv.release(); D.ds.nupd.release(); D.ds.rows_ptr.release();
  D.ds.upd_rows.release(); D.ds.rel.release(); D.ds.child_ptr.release(); D.ds.children.release();
  D.ds.asm_ptr.release(); D.ds.solver2v.release(); D.ds.lptr.release(); D.ds.uptr.release();
  D.ds.fbig.release(); D.ds.asm_entries.release(); D.ds.launch_list.release();
  D.ds.frow_ptr.release(); D.ds.gather_ptr.release(); D.ds.gather_src.release();
  D.tile_list.release(); D.tile_item_ptr.release(); D.tile_item_nv.release(); D.tile_fptr.release(); D.tile_items.release();
  D.est.release(); D.meas.release(); D.V.release(); D.Lv.release(); D.Uv.release(); D.uvec.release();
  D.x.release(); D.Fbig.release(); D.chi2.release(); D.chi2_part.release(); D.status.release();
  D.est0.release(); D.trig.release(); D.dbg_clocks.release(); D.timeline.release(); D.lmo_pose.release(); D.lmo_info.release();
  xchg_release(D);
  D.drop_graph();
  delete c->sys;
  c->sys = nullptr;
}

// assoc.cu -- cone data association on the device (sm_100a).
//
// Replaces, for whole frames at a time, Slam::coneToGlobal (slam.cpp:499-510),
// Slam::Spherical2Cartesian (637-654), Slam::transformConeToCoG (513-523),
// Slam::distanceBetweenCones (708-711), the matching/map-growth/loop-closure part of
// Slam::addConesToMap (552-623), Slam::loopClosing (697-706) and the matching loop of
// Slam::localizer (350-387).
//
// The reference association is FIRST-FIT in map-index order (575-607): the result for one
// observation is the MINIMUM map index among the gated candidates, which is what every kernel here
// computes (ascending scans that stop at the first hit, or an explicit min over bucket candidates).
//
// Compiled with -fmad=false: the reference build (x86-64 -O2, CMakeLists.txt:35-38) has no fused
// multiply-add, so products and sums are rounded separately exactly like the CPU does, and sqrt is
// IEEE correctly rounded on both sides: the gate arithmetic itself (coneToGlobal's rotation,
// distanceBetweenCones, the exact `d^2 < D*` test) is bit-identical to the reference's.
// What is NOT bit-identical is the polar -> Cartesian conversion in front of it: device sincos() /
// asin() are accurate to 1-2 ulp, glibc's libm (what the reference links) to < 1 ulp, and they do not
// round the same way.  The real guarantee therefore is: IDENTICAL association decisions unless an
// observation's distance to a candidate lies within a few ulp (relative 1e-15, i.e. ~1e-15 m) of
// sameConeThreshold or of the 1 m loop-closure gate; tests assert that no decision of any fixture
// sits within 1e-9 m of a gate (tests/test_assoc_gpu.py), and the converted coordinates themselves
// are checked at a 16-ulp bar against 424 vectors from the reference's compiled slam.cpp.
// The heading's cos/sin (coneToGlobal) are taken from the HOST's libm for exactly this reason.
#include <cub/device/device_scan.cuh>
#include <cmath>
#include <cfloat>

#include "ctx.h"

namespace {

__device__ __forceinline__ double ref_deg2rad() { return 0.017453292522222; }   // slam.hpp:134
__device__ __forceinline__ double ref_rad2deg() { return 57.295779513082325; }  // slam.hpp:135
__device__ __forceinline__ double ref_pi() { return (double)3.14159265f; }      // slam.hpp:136

// slam.cpp:513-523.  The angle leaves in degrees, as in the reference: the rad->deg->rad round trip
// rescales it by RAD2DEG*DEG2RAD != 1 (slam.hpp:134-135) and that is part of the result.  sin and cos
// of one argument come from one sincos() (one range reduction); `angle / fabs(angle)` is +-1 for
// every finite non-zero angle and NaN otherwise.
__device__ __forceinline__ void transform_cone_to_cog(double angle, double distance, double& angOut,
                                                      double& distOut) {
  const double lidarDistToCoG = 1.5;
  double sign = (angle != 0.0 && isfinite(angle)) ? copysign(1.0, angle) : __longlong_as_double(0x7ff8000000000000LL);
  angle = ref_pi() - fabs(angle * ref_deg2rad());
  double sa, ca;
  sincos(angle, &sa, &ca);
  double distanceNew = sqrt(lidarDistToCoG * lidarDistToCoG + distance * distance -
                            2 * lidarDistToCoG * distance * ca);
  double angleNew = asin((sa * distance) / distanceNew) * ref_rad2deg();
  angOut = angleNew * sign;
  distOut = distanceNew;
}

// slam.cpp:637-654.  zenith == 0 (every lidar cone of the 2-D pipeline): cos = 1 and sin = +-0
// exactly, so the products below are unchanged bit for bit and one sincos is saved.
__device__ __forceinline__ void spherical2cartesian(double az, double zen, double dist, double& x,
                                                    double& y, double& z) {
  double a, d;
  transform_cone_to_cog(az, dist, a, d);
  double sa, ca, sz, cz;
  sincos(a * ref_deg2rad(), &sa, &ca);
  const double zr = zen * ref_deg2rad();
  if (zr == 0.0) {
    sz = zr;
    cz = 1.0;
  } else {
    sincos(zr, &sz, &cz);
  }
  x = d * cz * ca;
  y = d * cz * sa;
  z = d * sz;
}

// slam.cpp:499-510 (pose trig hoisted by the callers: cp = cos(pose.theta), sp = sin(pose.theta))
__device__ __forceinline__ void local_to_global(double lx, double ly, double px, double py, double cp,
                                                double sp, double& gx, double& gy) {
  double newX = lx * cp - ly * sp;
  double newY = lx * sp + ly * cp;
  gx = newX + px;
  gy = newY + py;
}

// slam.cpp:708-711
__device__ __forceinline__ double cone_distance(double x1, double y1, double x2, double y2) {
  return sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
}
__device__ __forceinline__ double cone_distance2(double x1, double y1, double x2, double y2) {
  return (x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2);
}

// type gates.  mapping: fabs(mapType - obsType) < 0.0001 (slam.cpp:576), obsType the raw double.
// localizer: (mapType - (int)obsType) < 0.0001 (slam.cpp:358,360) == mapType <= (int)obsType.
template <int GATE>
__device__ __forceinline__ bool type_gate(int mapType, double obsType, int obsTypeInt) {
  if (GATE == SLAM_B200_GATE_MAPPING) return fabs((double)mapType - obsType) < 0.0001;
  return (mapType - obsTypeInt) < 0.0001;
}

// ------------------------------------------------------------------------------------------------
// frame kernels (one CTA; a frame is <= a few thousand observations, slam.cpp:46,244)
// ------------------------------------------------------------------------------------------------
constexpr int FRAME_THREADS = 512;
constexpr int FRAME_TILE = 2048;  // map cones staged in shared memory per pass (40 KB)

// Per-frame mailbox in mapped pinned host memory (slam_b200_assoc_map_frame / _localize_frame).  The reference's
// frames are ~8 columns against <= 300 cones: the kernel itself is a few microseconds, so two or three copy-engine
// operations and a stream synchronisation per frame (the first version: 50 us per call) cost more than the work.
// With a mailbox the kernel fetches its 4n + 5 input doubles from host memory once (one PCIe round trip), works in
// device memory as before, copies its records to the mailbox, and publishes the frame's sequence number behind a
// system-scope fence; the host spins on that word.  All pointers null = the copy path (SLAM_B200_FRAME_COPIES=1).
struct FrameMailbox {
  const double* in;   // 4n + 5 doubles
  int* outi;          // 2n + 8 ints
  double* outd;       // 5n doubles
  unsigned* flag;
  unsigned seq;
};
__device__ __forceinline__ void mailbox_fetch(const FrameMailbox& mb, double* in, int n) {
  if (!mb.in) return;
  for (int i = threadIdx.x; i < 4 * n + 5; i += blockDim.x) in[i] = mb.in[i];
  __syncthreads();
}
__device__ __forceinline__ void mailbox_publish(const FrameMailbox& mb, const int* outi, int ni, const double* outd, int nd) {
  if (!mb.in) return;
  __syncthreads();
  for (int i = threadIdx.x; i < ni; i += blockDim.x) mb.outi[i] = outi[i];
  for (int i = threadIdx.x; i < nd; i += blockDim.x) mb.outd[i] = outd[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(mb.flag) = mb.seq;
}

struct FrameScalars {  // ints at the head of the int output buffer
  int first_cone_created, loop_closing_obs, map_n, current_cone_index, loop_closing, n_reobserved,
      send_cone_data, pad;
};

// Phase 1 of both frame kernels: first-fit of every observation against map[0, M) (frozen),
// one warp per observation, map tiles staged in shared memory, ascending 32-wide chunks with a
// ballot so the first chunk containing a hit yields the minimum index.
template <int GATE>
__device__ void frame_first_fit(const double* __restrict__ gx, const double* __restrict__ gy,
                                const double* __restrict__ gt, int n, const double* map_x,
                                const double* map_y, const int* map_type, int M, double thr,
                                int* idx, double* sx, double* sy, int* st) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < n; i += blockDim.x) idx[i] = -1;
  for (int t0 = 0; t0 < M; t0 += FRAME_TILE) {
    int tn = min(FRAME_TILE, M - t0);
    __syncthreads();
    for (int k = threadIdx.x; k < tn; k += blockDim.x) {
      sx[k] = map_x[t0 + k];
      sy[k] = map_y[t0 + k];
      st[k] = map_type[t0 + k];
    }
    __syncthreads();
    for (int i = warp; i < n; i += nwarps) {
      if (idx[i] >= 0) continue;  // warp-uniform
      double ox = gx[i], oy = gy[i], ot = gt[i];
      int oti = (int)ot;
      int found = -1;
      for (int c = 0; c < tn && found < 0; c += 32) {
        int k = c + lane;
        bool hit = false;
        if (k < tn) hit = type_gate<GATE>(st[k], ot, oti) && cone_distance(sx[k], sy[k], ox, oy) < thr;
        unsigned b = __ballot_sync(0xffffffffu, hit);
        if (b) found = t0 + c + __ffs(b) - 1;
      }
      if (found >= 0 && lane == 0) idx[i] = found;
    }
  }
  __syncthreads();
}

// Mapping-phase frame (slam.cpp:552-623).  in: 4n frame doubles then pose(3).
__device__ __forceinline__ void map_frame_body(const double* in, int n, double thr, double mapThr, double* map_x,
                                               double* map_y, int* map_type, int M0, unsigned cci_in, int lc_in,
                                               double* outd, int* outi, double* sx, double* sy, int* st) {
  FrameScalars* sc = reinterpret_cast<FrameScalars*>(outi);
  int* idx = outi + 8;
  int* status = outi + 8 + n;
  double* z2 = outd;           // 2n
  double* g3 = outd + 2 * (size_t)n;  // 3n
  const double px = in[4 * (size_t)n], py = in[4 * (size_t)n + 1];
  const double cp = in[4 * (size_t)n + 3], sp = in[4 * (size_t)n + 4];
  // conversion of all columns (slam.cpp:572, 539)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double* col = in + 4 * (size_t)i;
    double lx, ly, lz, gx, gy;
    spherical2cartesian(col[0], col[1], col[2], lx, ly, lz);
    local_to_global(lx, ly, px, py, cp, sp, gx, gy);
    z2[2 * i] = lx;
    z2[2 * i + 1] = ly;
    g3[3 * i] = gx;
    g3[3 * i + 1] = gy;
    g3[3 * i + 2] = col[3];
  }
  __syncthreads();
  int M = M0;
  int first = 0;
  if (M == 0 && n > 0) {  // slam.cpp:554-567
    if (threadIdx.x == 0) {
      map_x[0] = g3[0];
      map_y[0] = g3[1];
      map_type[0] = (int)g3[2];
    }
    M = 1;
    first = 1;
    __syncthreads();
  }
  // SoA views of g for the shared routine: read straight from g3 with stride 3 via small lambdas
  // is awkward, so stage gx/gy/type contiguously behind g3 (outd has 8n doubles).
  double* gxs = outd + 5 * (size_t)n;
  double* gys = gxs + n;
  double* gts = gys + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    gxs[i] = g3[3 * i];
    gys[i] = g3[3 * i + 1];
    gts[i] = g3[3 * i + 2];
  }
  __syncthreads();
  if (!lc_in)
    frame_first_fit<SLAM_B200_GATE_MAPPING>(gxs, gys, gts, n, map_x, map_y, map_type, M, thr, idx, sx, sy, st);
  else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) idx[i] = -1;
    __syncthreads();
  }
  // Phase 2: the order-dependent part, one warp, observations in order: cones created earlier in
  // this frame are visible to later observations (611 inside the loop), m_currentConeIndex feeds
  // the loop-closure test (702), m_loopClosing short-circuits everything after it (575, 608).
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int Mpre = M;  // pre-frame map (incl. the first cone)
    unsigned cci = cci_in;
    int lc = lc_in, lcObs = -1;
    double minDistance = 100;  // 569
    double m0x = 0, m0y = 0;
    if (M > 0) { m0x = map_x[0]; m0y = map_y[0]; }
    for (int i = 0; i < n; i++) {
      double d2c = in[4 * (size_t)i + 2];
      double ox = gxs[i], oy = gys[i], ot = gts[i];
      int j = lc ? -1 : idx[i];
      if (!lc && j < 0) {
        for (int c = Mpre; c < M && j < 0; c += 32) {  // cones created earlier in this frame
          int k = c + lane;
          bool hit = false;
          if (k < M) {
            double kx = ((volatile double*)map_x)[k], ky = ((volatile double*)map_y)[k];
            int kt = ((volatile int*)map_type)[k];
            hit = type_gate<SLAM_B200_GATE_MAPPING>(kt, ot, 0) && cone_distance(kx, ky, ox, oy) < thr;
          }
          unsigned b = __ballot_sync(0xffffffffu, hit);
          if (b) j = c + __ffs(b) - 1;
        }
      }
      int stt;
      if (lc) {
        stt = SLAM_B200_ASSOC_SKIPPED;
        j = -1;
      } else if (j >= 0) {
        stt = SLAM_B200_ASSOC_MATCHED;
        double jx = ((volatile double*)map_x)[j], jy = ((volatile double*)map_y)[j];
        if (cone_distance(m0x, m0y, jx, jy) < 1 && cci > 20 && d2c < mapThr) {  // 697-706, 593-596
          lc = 1;
          lcObs = i;
        }
        if (d2c < minDistance) {  // 598-601
          cci = (unsigned)j;
          minDistance = d2c;
        }
      } else if (d2c < mapThr) {  // 608-611
        stt = SLAM_B200_ASSOC_NEW;
        j = M;
        if (lane == 0) {
          map_x[M] = ox;
          map_y[M] = oy;
          map_type[M] = (int)ot;
        }
        __threadfence_block();
        __syncwarp();
        if (M == 0) { m0x = ox; m0y = oy; }
        M++;
      } else {
        stt = SLAM_B200_ASSOC_NONE;
      }
      if (lane == 0) {
        idx[i] = j;
        status[i] = stt;
      }
    }
    if (lane == 0) {
      sc->first_cone_created = first;
      sc->loop_closing_obs = lcObs;
      sc->map_n = M;
      sc->current_cone_index = (int)cci;
      sc->loop_closing = lc;
      sc->n_reobserved = 0;
      sc->send_cone_data = 0;
      sc->pad = 0;
    }
  }
}

__global__ void __launch_bounds__(FRAME_THREADS, 1)
assoc_map_frame_kernel(double* in, int n, double thr, double mapThr, double* map_x,
                       double* map_y, int* map_type, int M0, unsigned cci_in, int lc_in,
                       double* outd, int* outi, FrameMailbox mb) {
  __shared__ double sx[FRAME_TILE], sy[FRAME_TILE];
  __shared__ int st[FRAME_TILE];
  mailbox_fetch(mb, in, n);
  map_frame_body(in, n, thr, mapThr, map_x, map_y, map_type, M0, cci_in, lc_in, outd, outi, sx, sy, st);
  mailbox_publish(mb, outi, 2 * n + 8, outd, 5 * n);
}

// Whole Monte-Carlo DRIVES per replica (SURVEY section 7 step 6, 8(d) "optional second variant" of config 3): one CTA
// per replica runs the mapping phase of performSLAM (slam.cpp:298-338 -> addConesToMap) frame after frame -- the body
// of the single-frame kernel above, unchanged -- with the replica's map in its own slice of the map arrays and the
// frame state (map size, m_currentConeIndex, m_loopClosing) carried from frame to frame on the device.  A replica
// stops at the frame that closes its loop (the reference then optimises and switches to the localiser); the records
// of every frame (scalars, idx, status -- the layout of slam_b200_assoc_map_frame) stay in HBM for the graph builder.
//   frames: [R][F][DRIVE_IN_STRIDE(nmax)] doubles: 4n cone values, pose(3), cos, sin at 4n (n = ncols[r][f])
__host__ __device__ inline size_t drive_in_stride(int nmax) { return 4 * (size_t)nmax + 8; }
__host__ __device__ inline size_t drive_rec_stride(int nmax) { return 2 * (size_t)nmax + 8; }

__global__ void __launch_bounds__(FRAME_THREADS, 2)
drive_replicas_kernel(const double* __restrict__ frames, const int* __restrict__ ncols, int F, int nmax, double thr,
                      double mapThr, double* map_x, double* map_y, int* map_type, int cap, double* scratch,
                      int* records, int* map_n_out, int* closed_at) {
  __shared__ double sx[FRAME_TILE], sy[FRAME_TILE];
  __shared__ int st[FRAME_TILE];
  const int r = blockIdx.x;
  double* mx = map_x + (size_t)r * cap;
  double* my = map_y + (size_t)r * cap;
  int* mt = map_type + (size_t)r * cap;
  double* outd = scratch + (size_t)r * (8 * (size_t)nmax + 8);
  int M = 0, lc = 0, closed = -1, overflow = 0;
  unsigned cci = 0;
  for (int f = 0; f < F; f++) {
    const int n = ncols[(size_t)r * F + f];
    int* outi = records + ((size_t)r * F + f) * drive_rec_stride(nmax);
    const bool run = !lc && !overflow && n > 0 && n <= nmax && M + n + 1 <= cap;
    if (!run) {  // block-uniform: the frame is recorded as not run (n_reobserved = -1)
      if (!lc && n > 0 && (n > nmax || M + n + 1 > cap)) overflow = 1;
      if (threadIdx.x == 0) {
        FrameScalars* sc = reinterpret_cast<FrameScalars*>(outi);
        sc->first_cone_created = 0; sc->loop_closing_obs = -1; sc->map_n = M; sc->current_cone_index = (int)cci;
        sc->loop_closing = lc; sc->n_reobserved = -1; sc->send_cone_data = overflow; sc->pad = 0;
      }
      for (int i = threadIdx.x; i < min(n, nmax); i += blockDim.x) {
        outi[8 + i] = -1;
        outi[8 + min(n, nmax) + i] = SLAM_B200_ASSOC_SKIPPED;
      }
      continue;
    }
    const double* in = frames + ((size_t)r * F + f) * drive_in_stride(nmax);
    map_frame_body(in, n, thr, mapThr, mx, my, mt, M, cci, lc, outd, outi, sx, sy, st);
    __syncthreads();  // the frame's scalars (written by lane 0) are the next frame's state
    const FrameScalars* sc = reinterpret_cast<const FrameScalars*>(outi);
    M = sc->map_n;
    cci = (unsigned)sc->current_cone_index;
    lc = sc->loop_closing;
    if (lc && closed < 0) closed = f;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    map_n_out[r] = M;
    closed_at[r] = overflow ? -2 : closed;
  }
}

// Localisation-phase frame (slam.cpp:350-387)
__global__ void __launch_bounds__(FRAME_THREADS, 1)
assoc_localize_frame_kernel(double* in, int n, double thr, const double* map_x,
                            const double* map_y, const int* map_type, int M, unsigned cci_in,
                            double* outd, int* outi, FrameMailbox mb) {
  __shared__ double sx[FRAME_TILE], sy[FRAME_TILE];
  __shared__ int st[FRAME_TILE];
  mailbox_fetch(mb, in, n);
  FrameScalars* sc = reinterpret_cast<FrameScalars*>(outi);
  int* idx = outi + 8;
  double* g3 = outd + 2 * (size_t)n;
  double* gxs = outd + 5 * (size_t)n;
  double* gys = gxs + n;
  double* gts = gys + n;
  const double px = in[4 * (size_t)n], py = in[4 * (size_t)n + 1];
  const double cp = in[4 * (size_t)n + 3], sp = in[4 * (size_t)n + 4];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double* col = in + 4 * (size_t)i;
    double lx, ly, lz, gx, gy;
    spherical2cartesian(col[0], col[1], col[2], lx, ly, lz);
    local_to_global(lx, ly, px, py, cp, sp, gx, gy);
    g3[3 * i] = gx; g3[3 * i + 1] = gy; g3[3 * i + 2] = col[3];
    gxs[i] = gx; gys[i] = gy; gts[i] = col[3];
  }
  __syncthreads();
  frame_first_fit<SLAM_B200_GATE_LOCALIZER>(gxs, gys, gts, n, map_x, map_y, map_type, M, thr, idx, sx, sy, st);
  if (threadIdx.x == 0) {  // 375-378, 385-387: nearest matched observation, strict <, first wins
    double minDistance = 100;
    unsigned cci = 0;
    int reobs = 0;
    for (int i = 0; i < n; i++) {
      if (idx[i] >= 0) {
        reobs++;
        double d2c = in[4 * (size_t)i + 2];
        if (d2c < minDistance) { cci = (unsigned)idx[i]; minDistance = d2c; }
      }
    }
    sc->first_cone_created = 0;
    sc->loop_closing_obs = -1;
    sc->map_n = M;
    sc->n_reobserved = reobs;
    sc->send_cone_data = reobs > 0 ? (cci != cci_in) : 0;
    sc->current_cone_index = reobs > 0 ? (int)cci : (int)cci_in;
    sc->loop_closing = 0;
    sc->pad = 0;
  }
  mailbox_publish(mb, outi, n + 8, outd, 5 * n);
}

// Slam::updateMap on the device (slam.cpp:713-732): map cone j <- estimate of its landmark vertex.
// est = x[P] | y[P] | theta[P] | lx[L] | ly[L] (graph_dev.h)
__global__ void map_from_graph_kernel(const double* __restrict__ est, int P, int L, const int* __restrict__ lm_of_map,
                                      double* map_x, double* map_y, int M) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int l = lm_of_map[j];
  if (l < 0) return;
  map_x[j] = est[3 * (size_t)P + l];
  map_y[j] = est[3 * (size_t)P + L + l];
}

// conversion only (slam_b200_cones_to_global)
__global__ void convert_kernel(const double* __restrict__ in, int n, double* __restrict__ g3,
                               double* __restrict__ l3) {
  const double px = in[4 * (size_t)n], py = in[4 * (size_t)n + 1];
  const double cp = in[4 * (size_t)n + 3], sp = in[4 * (size_t)n + 4];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* col = in + 4 * (size_t)i;
  double lx, ly, lz, gx, gy;
  spherical2cartesian(col[0], col[1], col[2], lx, ly, lz);
  local_to_global(lx, ly, px, py, cp, sp, gx, gy);
  g3[3 * i] = gx; g3[3 * i + 1] = gy; g3[3 * i + 2] = col[3];
  l3[3 * i] = lx; l3[3 * i + 1] = ly; l3[3 * i + 2] = lz;
}

// ------------------------------------------------------------------------------------------------
// bulk match-only kernels (large cone field).  thr2x = smallest double s with sqrt(s) >= thr, so
// `d2 < thr2x` decides exactly like `sqrt(d2) < thr` (sqrt is monotone and correctly rounded)
// without a square root in the inner loop.
// ------------------------------------------------------------------------------------------------
constexpr int BULK_THREADS = 128;
constexpr int BULK_TILE = 1024;

struct PoseTrig { double px, py, cp, sp; };

__device__ __forceinline__ void load_obs(const double* __restrict__ cones, int i, const PoseTrig& pt,
                                         double& gx, double& gy, double& ot) {
  // one observation = one 32-byte column: two 16-byte loads
  const double2* c2 = reinterpret_cast<const double2*>(cones + 4 * (size_t)i);
  double2 a = __ldg(c2), b = __ldg(c2 + 1);
  double lx, ly, lz;
  spherical2cartesian(a.x, a.y, b.x, lx, ly, lz);
  local_to_global(lx, ly, pt.px, pt.py, pt.cp, pt.sp, gx, gy);
  ot = b.y;
}

// brute force: one thread per observation, the whole map streamed through shared memory in tiles,
// ascending index, stop at the first hit.
template <int GATE>
__global__ void __launch_bounds__(BULK_THREADS)
assoc_bulk_brute_kernel(const double* __restrict__ cones, int n, PoseTrig pt, double thr2x,
                        const double* __restrict__ map_x, const double* __restrict__ map_y,
                        const int* __restrict__ map_type, int M, int* __restrict__ idx) {
  __shared__ double sx[BULK_TILE], sy[BULK_TILE];
  __shared__ int st[BULK_TILE];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double gx = 0, gy = 0, ot = 0;
  int found = -1;
  bool active = i < n;
  if (active) load_obs(cones, i, pt, gx, gy, ot);
  int oti = (int)ot;
  for (int t0 = 0; t0 < M; t0 += BULK_TILE) {
    int tn = min(BULK_TILE, M - t0);
    if (__syncthreads_and(!active || found >= 0)) break;
    for (int k = threadIdx.x; k < tn; k += blockDim.x) {
      sx[k] = __ldg(map_x + t0 + k);
      sy[k] = __ldg(map_y + t0 + k);
      st[k] = __ldg(map_type + t0 + k);
    }
    __syncthreads();
    if (active && found < 0) {
      for (int k = 0; k < tn; k++) {
        if (type_gate<GATE>(st[k], ot, oti) && cone_distance2(sx[k], sy[k], gx, gy) < thr2x) {
          found = t0 + k;
          break;
        }
      }
    }
  }
  if (active) idx[i] = found;
}

// ---- uniform grid index --------------------------------------------------------------------------
__global__ void bbox_kernel(const double* __restrict__ x, const double* __restrict__ y, int M,
                            double* __restrict__ bbox) {  // bbox pre-set to +inf,+inf,-inf,-inf
  double mnx = DBL_MAX, mny = DBL_MAX, mxx = -DBL_MAX, mxy = -DBL_MAX;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    double a = x[i], b = y[i];
    if (isfinite(a) && isfinite(b)) {
      mnx = fmin(mnx, a); mxx = fmax(mxx, a);
      mny = fmin(mny, b); mxy = fmax(mxy, b);
    }
  }
  for (int o = 16; o; o >>= 1) {
    mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if ((threadIdx.x & 31) == 0) {
    // doubles are ordered like their sign-magnitude bit patterns; use CAS loops on the 4 slots
    auto amin = [](double* addr, double v) {
      unsigned long long* a = (unsigned long long*)addr;
      unsigned long long old = *a, assumed;
      do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) <= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
      } while (assumed != old);
    };
    auto amax = [](double* addr, double v) {
      unsigned long long* a = (unsigned long long*)addr;
      unsigned long long old = *a, assumed;
      do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
      } while (assumed != old);
    };
    amin(bbox + 0, mnx); amin(bbox + 1, mny); amax(bbox + 2, mxx); amax(bbox + 3, mxy);
  }
}

struct GridParams { double x0, y0, inv, h; int nx, ny; };

__device__ __forceinline__ int cell_coord(double v, double v0, double inv) {
  // floor((v - v0) * inv): monotone non-decreasing in v (fp subtract and multiply by a positive
  // constant are monotone), which is all the bucket argument needs
  return (int)floor((v - v0) * inv);
}

__global__ void grid_count_kernel(const double* __restrict__ x, const double* __restrict__ y, int M,
                                  GridParams gp, int* __restrict__ count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double a = x[i], b = y[i];
  if (!(isfinite(a) && isfinite(b))) return;  // a NaN cone can never pass the distance gate
  int cx = min(max(cell_coord(a, gp.x0, gp.inv), 0), gp.nx - 1);
  int cy = min(max(cell_coord(b, gp.y0, gp.inv), 0), gp.ny - 1);
  atomicAdd(count + (size_t)cy * gp.nx + cx, 1);
}

__global__ void grid_fill_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                 const int* __restrict__ type, int M, GridParams gp,
                                 int* __restrict__ cursor, GridRec* __restrict__ rec) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  double a = x[i], b = y[i];
  if (!(isfinite(a) && isfinite(b))) return;
  int cx = min(max(cell_coord(a, gp.x0, gp.inv), 0), gp.nx - 1);
  int cy = min(max(cell_coord(b, gp.y0, gp.inv), 0), gp.ny - 1);
  int slot = atomicAdd(cursor + (size_t)cy * gp.nx + cx, 1);
  GridRec r;
  r.x = a; r.y = b; r.type = type[i]; r.idx = i; r.pad0 = 0; r.pad1 = 0;
  rec[slot] = r;
}

// bucketed: one thread per observation.  Cells are 2*h wide (h = threshold plus a 1e-9 relative
// margin), so the disc of radius thr around the observation overlaps at most 2 x 2 cells = two
// contiguous runs of the cell-sorted map.  Every cone is one aligned 32-byte record (one DRAM
// sector); both runs are walked as one sequence whose records are fetched GRID_BATCH at a time before
// any is tested, so the dependent chain per observation is: observation -> cell table -> records.  The answer is the minimum
// ORIGINAL index over the gated candidates (= the reference's first fit, slam.cpp:575-607).
constexpr int GRID_BATCH = 8;  // records fetched before any is tested

//
// EARLY (SLAM_B200_ALGO_GRID_PIPELINED): the kernel is latency-bound -- one wave of threads, each
// a chain of three dependent memory round trips -- and leaves most of the machine idle, so frames
// against the frozen map are launched with programmatic stream serialisation and every CTA
// releases the next frame's launch at once (griddepcontrol.launch_dependents, no
// griddepcontrol.wait: the frames are independent); successive frames overlap on one stream.
template <int GATE>
__device__ __forceinline__ void grid_match_one(const double* __restrict__ cones, int i, const PoseTrig& pt, double thr2x,
                                               const GridParams& gp, const int* __restrict__ cell_start,
                                               const GridRec* __restrict__ rec, int* __restrict__ idx) {
  double gx, gy, ot;
  load_obs(cones, i, pt, gx, gy, ot);
  int oti = (int)ot;
  int best = 0x7fffffff;
  if (isfinite(gx) && isfinite(gy)) {
    const double fx0 = floor((gx - gp.h - gp.x0) * gp.inv), fx1 = floor((gx + gp.h - gp.x0) * gp.inv);
    const double fy0 = floor((gy - gp.h - gp.y0) * gp.inv), fy1 = floor((gy + gp.h - gp.y0) * gp.inv);
    // observations whose disc misses the map's bounding box have no candidates
    if (fx1 >= 0.0 && fx0 <= (double)(gp.nx - 1) && fy1 >= 0.0 && fy0 <= (double)(gp.ny - 1)) {
      const int cx0 = (int)fmax(fx0, 0.0), cx1 = (int)fmin(fx1, (double)(gp.nx - 1));
      const int cy0 = (int)fmax(fy0, 0.0), cy1 = (int)fmin(fy1, (double)(gp.ny - 1));
      const size_t b0 = (size_t)cy0 * gp.nx, b1 = (size_t)cy1 * gp.nx;
      // four independent cell-table reads (the second row repeats the first when the disc stays
      // inside one row of cells; its run is then emptied)
      const int s0 = __ldg(cell_start + b0 + cx0), e0 = __ldg(cell_start + b0 + cx1 + 1);
      const int s1 = __ldg(cell_start + b1 + cx0);
      const int e1 = cy1 > cy0 ? __ldg(cell_start + b1 + cx1 + 1) : s1;
      // the two runs are walked as one sequence, GRID_BATCH records in flight at a time
      const int len0 = e0 - s0, len = len0 + (e1 - s1);
      for (int k = 0; k < len; k += GRID_BATCH) {
        double2 xy[GRID_BATCH];
        int2 ti[GRID_BATCH];
#pragma unroll
        for (int q = 0; q < GRID_BATCH; q++) {
          const int kk = k + q;
          if (kk < len) {
            const GridRec* r = rec + (kk < len0 ? s0 + kk : s1 + (kk - len0));
            xy[q] = __ldg(reinterpret_cast<const double2*>(r));
            ti[q] = __ldg(reinterpret_cast<const int2*>(r) + 2);
          }
        }
#pragma unroll
        for (int q = 0; q < GRID_BATCH; q++) {
          if (k + q < len) {
            if (type_gate<GATE>(ti[q].x, ot, oti) && ti[q].y < best &&
                cone_distance2(xy[q].x, xy[q].y, gx, gy) < thr2x)
              best = ti[q].y;
          }
        }
      }
    }
  }
  idx[i] = best == 0x7fffffff ? -1 : best;
}

template <int GATE, bool EARLY>
__global__ void __launch_bounds__(BULK_THREADS)
assoc_bulk_grid_kernel(const double* __restrict__ cones, int n, PoseTrig pt, double thr2x, GridParams gp,
                       const int* __restrict__ cell_start, const GridRec* __restrict__ rec,
                       int* __restrict__ idx) {
  if (EARLY) asm volatile("griddepcontrol.launch_dependents;");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  grid_match_one<GATE>(cones, i, pt, thr2x, gp, cell_start, rec, idx);
}

// SLAM_B200_ALGO_GRID_BATCHED: up to BATCH_FRAMES (32) independent frames against frozen maps in ONE launch
// (blockIdx.y = frame).  A single frame is one sub-wave of threads on a three-deep dependent-load chain and cannot
// fill the memory system however it is written; several frames in one grid can, without
// the per-launch cost the pipelined train still pays for every frame (about 2 us each: what caps a rank that holds
// an eighth of the observations).
constexpr int BATCH_FRAMES = 32;  // 32 x 112 bytes of per-frame parameters: inside the 4 KB kernel-parameter block
struct BatchFrame {
  const double* cones;
  const int* cell_start;
  const GridRec* rec;
  int* idx;
  PoseTrig pt;
  GridParams gp;
  int n;
};
struct BatchParams { BatchFrame f[BATCH_FRAMES]; };

template <int GATE>
__global__ void __launch_bounds__(BULK_THREADS)
assoc_bulk_grid_batched_kernel(const __grid_constant__ BatchParams bp, double thr2x) {
  const BatchFrame& F = bp.f[blockIdx.y];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F.n) return;
  grid_match_one<GATE>(F.cones, i, F.pt, thr2x, F.gp, F.cell_start, F.rec, F.idx);
}

// smallest double s with sqrt(s) >= thr  (host; IEEE sqrt on both sides)
double sqrt_gate_threshold(double thr) {
  if (!(thr > 0)) return 0.0;  // dist < thr never true for thr <= 0 (dist >= 0)
  double s = thr * thr;
  while (std::sqrt(s) >= thr) s = std::nextafter(s, -INFINITY);
  while (std::sqrt(s) < thr) s = std::nextafter(s, INFINITY);
  return s;
}

int ensure_frame_buffers(slam_b200_ctx* c, int n) {
  SLAM_CUDA_TRY(c, c->frame_in.exact(4 * (size_t)n + 8));
  SLAM_CUDA_TRY(c, c->frame_outd.exact(8 * (size_t)n + 8));
  SLAM_CUDA_TRY(c, c->frame_outi.exact(2 * (size_t)n + 8));
  SLAM_CUDA_TRY(c, c->pin_d.reserve(8 * (size_t)n + 8));
  SLAM_CUDA_TRY(c, c->pin_i.reserve(2 * (size_t)n + 8));
  return 0;
}

int upload_frame(slam_b200_ctx* c, const double* cones, int n, const double pose[3]) {
  std::memcpy(c->pin_d.p, cones, sizeof(double) * 4 * (size_t)n);
  std::memcpy(c->pin_d.p + 4 * (size_t)n, pose, sizeof(double) * 3);
  // cos/sin of the heading are two scalars per frame: taken from the host libm the reference
  // itself uses (slam.cpp:504-505) so every column is rotated by bit-identical factors
  c->pin_d.p[4 * (size_t)n + 3] = std::cos(pose[2]);
  c->pin_d.p[4 * (size_t)n + 4] = std::sin(pose[2]);
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->frame_in.p, c->pin_d.p, sizeof(double) * (4 * (size_t)n + 5),
                                   cudaMemcpyHostToDevice, c->stream));
  return 0;
}

// ---- mailbox (host side) ----
bool frame_copies() {
  static const bool v = getenv("SLAM_B200_FRAME_COPIES") != nullptr;
  return v;
}
inline size_t mbox_off_in() { return 64; }
inline size_t mbox_off_outi(size_t cap) { return mbox_off_in() + sizeof(double) * (4 * cap + 8); }
inline size_t mbox_off_outd(size_t cap) { return mbox_off_outi(cap) + sizeof(int) * (2 * cap + 8); }
inline size_t mbox_bytes(size_t cap) { return mbox_off_outd(cap) + sizeof(double) * (5 * cap + 8); }

int mailbox_reserve(slam_b200_ctx* c, int n) {
  if (c->mbox_h && (size_t)n <= c->mbox_cap) return 0;
  size_t cap = c->mbox_cap ? c->mbox_cap : 64;
  while (cap < (size_t)n) cap *= 2;
  if (c->mbox_h) {
    SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    cudaFreeHost(c->mbox_h);
    c->mbox_h = c->mbox_d = nullptr;
    c->mbox_cap = 0;
  }
  void *h = nullptr, *d = nullptr;
  SLAM_CUDA_TRY(c, cudaHostAlloc(&h, mbox_bytes(cap), cudaHostAllocMapped));
  cudaError_t e = cudaHostGetDevicePointer(&d, h, 0);
  if (e != cudaSuccess) { cudaFreeHost(h); SLAM_CUDA_TRY(c, e); }
  std::memset(h, 0, mbox_bytes(cap));
  c->mbox_h = static_cast<char*>(h);
  c->mbox_d = static_cast<char*>(d);
  c->mbox_cap = cap;
  c->mbox_seq = 0;
  return 0;
}

// writes the frame into the mailbox and returns the kernel-side view of it
FrameMailbox mailbox_fill(slam_b200_ctx* c, const double* cones, int n, const double pose[3]) {
  double* in = reinterpret_cast<double*>(c->mbox_h + mbox_off_in());
  std::memcpy(in, cones, sizeof(double) * 4 * (size_t)n);
  std::memcpy(in + 4 * (size_t)n, pose, sizeof(double) * 3);
  in[4 * (size_t)n + 3] = std::cos(pose[2]);  // host libm, like upload_frame
  in[4 * (size_t)n + 4] = std::sin(pose[2]);
  FrameMailbox mb;
  mb.in = reinterpret_cast<const double*>(c->mbox_d + mbox_off_in());
  mb.outi = reinterpret_cast<int*>(c->mbox_d + mbox_off_outi(c->mbox_cap));
  mb.outd = reinterpret_cast<double*>(c->mbox_d + mbox_off_outd(c->mbox_cap));
  mb.flag = reinterpret_cast<unsigned*>(c->mbox_d);
  mb.seq = ++c->mbox_seq;
  if (mb.seq == 0) mb.seq = ++c->mbox_seq;  // 0 is the cleared word
  return mb;
}

// spins until the kernel has published mb.seq; a kernel that ended without publishing (a fault) is reported
int mailbox_wait(slam_b200_ctx* c, unsigned seq) {
  volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(c->mbox_h);
  for (unsigned long spins = 1;; spins++) {
    if (*flag == seq) break;
    if ((spins & 0xfff) == 0) {
      cudaError_t q = cudaStreamQuery(c->stream);
      if (q != cudaErrorNotReady) {
        if (*flag == seq) break;
        if (q == cudaSuccess) { c->fail("frame kernel ended without publishing its records"); return SLAM_B200_E_CUDA; }
        SLAM_CUDA_TRY(c, q);
      }
    }
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  return 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int slam_b200_cones_to_global(slam_b200_ctx* c, const double* cones, int n, const double pose[3],
                              double* global3, double* local3) try {
  NvtxRange nvtx_range("slam_b200/cones_to_global");
  if (!c || !cones || !pose || n < 0) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  if (n == 0) return 0;
  if (int rc = ensure_frame_buffers(c, n)) return rc;
  if (int rc = upload_frame(c, cones, n, pose)) return rc;
  double* g3 = c->frame_outd.p;
  double* l3 = c->frame_outd.p + 3 * (size_t)n;
  convert_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->frame_in.p, n, g3, l3);
  c->launches++;
  SLAM_CUDA_TRY(c, cudaGetLastError());
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->pin_d.p, c->frame_outd.p, sizeof(double) * 6 * (size_t)n,
                                   cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (global3) std::memcpy(global3, c->pin_d.p, sizeof(double) * 3 * (size_t)n);
  if (local3) std::memcpy(local3, c->pin_d.p + 3 * (size_t)n, sizeof(double) * 3 * (size_t)n);
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_map_clear(slam_b200_ctx* c) try {
  if (!c) return SLAM_B200_E_ARG;
  c->map_n = 0;
  c->map_version++;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_map_append(slam_b200_ctx* c, const double* x, const double* y, const int32_t* type, int n) try {
  if (!c || n < 0 || (n > 0 && (!x || !y || !type))) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  if (n == 0) return c->map_n;
  size_t need = (size_t)c->map_n + n;
  SLAM_CUDA_TRY(c, c->map_x.reserve(need, c->map_n, c->stream));
  SLAM_CUDA_TRY(c, c->map_y.reserve(need, c->map_n, c->stream));
  SLAM_CUDA_TRY(c, c->map_type.reserve(need, c->map_n, c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->map_x.p + c->map_n, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->map_y.p + c->map_n, y, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->map_type.p + c->map_n, type, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->map_n += n;
  c->map_version++;
  return c->map_n;
} SLAM_ABI_CATCH(c)

int slam_b200_map_size(slam_b200_ctx* c) { return c ? c->map_n : SLAM_B200_E_ARG; }

int slam_b200_map_read(slam_b200_ctx* c, int first, int n, double* x, double* y, int32_t* type) try {
  if (!c || first < 0 || n < 0 || first + n > c->map_n) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  if (n == 0) return 0;
  if (x) SLAM_CUDA_TRY(c, cudaMemcpyAsync(x, c->map_x.p + first, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  if (y) SLAM_CUDA_TRY(c, cudaMemcpyAsync(y, c->map_y.p + first, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  if (type) SLAM_CUDA_TRY(c, cudaMemcpyAsync(type, c->map_type.p + first, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return n;
} SLAM_ABI_CATCH(c)

int slam_b200_map_write_xy(slam_b200_ctx* c, int first, int n, const double* x, const double* y) try {
  if (!c || first < 0 || n < 0 || first + n > c->map_n || (n > 0 && (!x || !y))) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  if (n == 0) return 0;
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->map_x.p + first, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->map_y.p + first, y, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->map_version++;
  return n;
} SLAM_ABI_CATCH(c)

// enqueues the copy of the device map into the pinned mirror and records the event readers wait for
static int mirror_refresh_async(slam_b200_ctx* c) {
  const size_t M = (size_t)c->map_n;
  if (M > c->mirror_cap) {
    if (c->mirror_event) SLAM_CUDA_TRY(c, cudaEventSynchronize(c->mirror_event));
    size_t cap = c->mirror_cap ? c->mirror_cap : 512;
    while (cap < M) cap *= 2;
    SLAM_CUDA_TRY(c, c->mirror_xy.reserve(2 * cap));
    SLAM_CUDA_TRY(c, c->mirror_type.reserve(cap));
    c->mirror_cap = cap;
  }
  if (!c->mirror_event) SLAM_CUDA_TRY(c, cudaEventCreateWithFlags(&c->mirror_event, cudaEventDisableTiming));
  if (M) {
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->mirror_xy.p, c->map_x.p, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->mirror_xy.p + c->mirror_cap, c->map_y.p, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->mirror_type.p, c->map_type.p, sizeof(int) * M, cudaMemcpyDeviceToHost, c->stream));
  }
  SLAM_CUDA_TRY(c, cudaEventRecord(c->mirror_event, c->stream));
  c->mirror_n = (int)M;
  c->mirror_version = c->map_version;
  return 0;
}

int slam_b200_map_update_from_graph(slam_b200_ctx* c) try {
  NvtxRange nvtx_range("slam_b200/map_update_from_graph");
  if (!c) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  const int M = c->map_n;
  if (M == 0) return 0;
  HostGraph& g = c->g;
  // landmark behind every map cone: the reference numbers landmark vertices by map index (slam.cpp:556,610)
  std::vector<int> lm(M, -1);
  int found = 0;
  for (int j = 0; j < M; j++) {
    const int v = g.id2v.get(j);
    if (v >= 0 && (v & 1)) { lm[j] = v >> 1; found++; }
  }
  if (found == 0) return 0;
  double *est_dev = nullptr;
  int P = 0, L = 0;
  if (graph_device_estimates(c, &est_dev, &P, &L)) {
    // device estimates are current (the usual case: right after an optimise): cone <- vertex on the device
    SLAM_CUDA_TRY(c, c->lm_of_map.exact((size_t)M));
    SLAM_CUDA_TRY(c, c->pin_i.reserve((size_t)M));
    std::memcpy(c->pin_i.p, lm.data(), sizeof(int) * (size_t)M);
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->lm_of_map.p, c->pin_i.p, sizeof(int) * (size_t)M, cudaMemcpyHostToDevice, c->stream));
    map_from_graph_kernel<<<(M + 255) / 256, 256, 0, c->stream>>>(est_dev, P, L, c->lm_of_map.p, c->map_x.p, c->map_y.p, M);
    c->launches++;
    SLAM_CUDA_TRY(c, cudaGetLastError());
  } else {
    // the host graph holds newer values than the device (set_values / add_* since the last optimise): write them
    std::vector<double> x(M), y(M);
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(x.data(), c->map_x.p, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(y.data(), c->map_y.p, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
    SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int j = 0; j < M; j++)
      if (lm[j] >= 0) { x[j] = g.lm_est[2 * (size_t)lm[j]]; y[j] = g.lm_est[2 * (size_t)lm[j] + 1]; }
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->map_x.p, x.data(), sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->map_y.p, y.data(), sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  }
  c->map_version++;
  if (int rc = mirror_refresh_async(c)) return rc;
  return found;
} SLAM_ABI_CATCH(c)

int slam_b200_map_mirror(slam_b200_ctx* c, const double** x, const double** y, const int32_t** type) try {
  if (!c) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  if (c->mirror_version != c->map_version || c->mirror_n != c->map_n)
    if (int rc = mirror_refresh_async(c)) return rc;
  if (c->mirror_event) SLAM_CUDA_TRY(c, cudaEventSynchronize(c->mirror_event));
  if (x) *x = c->mirror_xy.p;
  if (y) *y = c->mirror_xy.p + c->mirror_cap;
  if (type) *type = c->mirror_type.p;
  return c->mirror_n;
} SLAM_ABI_CATCH(c)

int slam_b200_assoc_map_frame(slam_b200_ctx* c, const double* cones, int n, const double pose[3],
                              double thr, double mapThr, uint32_t* cci, int32_t* loop_closing,
                              int32_t* idx, int32_t* status, double* z2, double* g3,
                              int32_t* first_cone_created, int32_t* loop_closing_obs) try {
  NvtxRange nvtx_range("slam_b200/assoc_map_frame");
  if (!c || !pose || !cci || !loop_closing || n < 0 || (n > 0 && (!cones || !idx || !status)))
    return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  if (first_cone_created) *first_cone_created = 0;
  if (loop_closing_obs) *loop_closing_obs = -1;
  if (n == 0) return c->map_n;
  if (int rc = ensure_frame_buffers(c, n)) return rc;
  size_t need = (size_t)c->map_n + n + 1;  // every column may create a cone
  SLAM_CUDA_TRY(c, c->map_x.reserve(need, c->map_n, c->stream));
  SLAM_CUDA_TRY(c, c->map_y.reserve(need, c->map_n, c->stream));
  SLAM_CUDA_TRY(c, c->map_type.reserve(need, c->map_n, c->stream));
  const int* ri;      // records: ints (8 scalars, idx, status) and doubles (z2, g3)
  const double* rd;
  if (!frame_copies()) {
    if (int rc = mailbox_reserve(c, n)) return rc;
    const FrameMailbox mb = mailbox_fill(c, cones, n, pose);
    assoc_map_frame_kernel<<<1, FRAME_THREADS, 0, c->stream>>>(c->frame_in.p, n, thr, mapThr, c->map_x.p,
                                                              c->map_y.p, c->map_type.p, c->map_n, *cci,
                                                              *loop_closing, c->frame_outd.p, c->frame_outi.p, mb);
    c->launches++;
    SLAM_CUDA_TRY(c, cudaGetLastError());
    if (int rc = mailbox_wait(c, mb.seq)) return rc;
    ri = reinterpret_cast<const int*>(c->mbox_h + mbox_off_outi(c->mbox_cap));
    rd = reinterpret_cast<const double*>(c->mbox_h + mbox_off_outd(c->mbox_cap));
  } else {
    if (int rc = upload_frame(c, cones, n, pose)) return rc;
    assoc_map_frame_kernel<<<1, FRAME_THREADS, 0, c->stream>>>(c->frame_in.p, n, thr, mapThr, c->map_x.p,
                                                              c->map_y.p, c->map_type.p, c->map_n, *cci,
                                                              *loop_closing, c->frame_outd.p, c->frame_outi.p,
                                                              FrameMailbox{nullptr, nullptr, nullptr, nullptr, 0});
    c->launches++;
    SLAM_CUDA_TRY(c, cudaGetLastError());
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->pin_i.p, c->frame_outi.p, sizeof(int) * (2 * (size_t)n + 8),
                                     cudaMemcpyDeviceToHost, c->stream));
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->pin_d.p, c->frame_outd.p, sizeof(double) * 5 * (size_t)n,
                                     cudaMemcpyDeviceToHost, c->stream));
    SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ri = c->pin_i.p;
    rd = c->pin_d.p;
  }
  const FrameScalars* sc = reinterpret_cast<const FrameScalars*>(ri);
  std::memcpy(idx, ri + 8, sizeof(int) * n);
  std::memcpy(status, ri + 8 + n, sizeof(int) * n);
  if (z2) std::memcpy(z2, rd, sizeof(double) * 2 * (size_t)n);
  if (g3) std::memcpy(g3, rd + 2 * (size_t)n, sizeof(double) * 3 * (size_t)n);
  if (first_cone_created) *first_cone_created = sc->first_cone_created;
  if (loop_closing_obs) *loop_closing_obs = sc->loop_closing_obs;
  *cci = (uint32_t)sc->current_cone_index;
  *loop_closing = sc->loop_closing;
  if (sc->map_n != c->map_n) c->map_version++;
  c->map_n = sc->map_n;
  return c->map_n;
} SLAM_ABI_CATCH(c)

int slam_b200_assoc_localize_frame(slam_b200_ctx* c, const double* cones, int n, const double pose[3],
                                   double thr, uint32_t* cci, int32_t* idx, double* g3,
                                   int32_t* n_reobserved, int32_t* send_cone_data) try {
  NvtxRange nvtx_range("slam_b200/assoc_localize_frame");
  if (!c || !pose || !cci || n < 0 || (n > 0 && (!cones || !idx))) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  if (n_reobserved) *n_reobserved = 0;
  if (send_cone_data) *send_cone_data = 0;
  if (n == 0) return 0;
  if (int rc = ensure_frame_buffers(c, n)) return rc;
  const int* ri;
  const double* rg3;
  if (!frame_copies()) {
    if (int rc = mailbox_reserve(c, n)) return rc;
    const FrameMailbox mb = mailbox_fill(c, cones, n, pose);
    assoc_localize_frame_kernel<<<1, FRAME_THREADS, 0, c->stream>>>(c->frame_in.p, n, thr, c->map_x.p, c->map_y.p,
                                                                   c->map_type.p, c->map_n, *cci,
                                                                   c->frame_outd.p, c->frame_outi.p, mb);
    c->launches++;
    SLAM_CUDA_TRY(c, cudaGetLastError());
    if (int rc = mailbox_wait(c, mb.seq)) return rc;
    ri = reinterpret_cast<const int*>(c->mbox_h + mbox_off_outi(c->mbox_cap));
    rg3 = reinterpret_cast<const double*>(c->mbox_h + mbox_off_outd(c->mbox_cap)) + 2 * (size_t)n;
  } else {
    if (int rc = upload_frame(c, cones, n, pose)) return rc;
    assoc_localize_frame_kernel<<<1, FRAME_THREADS, 0, c->stream>>>(c->frame_in.p, n, thr, c->map_x.p, c->map_y.p,
                                                                   c->map_type.p, c->map_n, *cci,
                                                                   c->frame_outd.p, c->frame_outi.p,
                                                                   FrameMailbox{nullptr, nullptr, nullptr, nullptr, 0});
    c->launches++;
    SLAM_CUDA_TRY(c, cudaGetLastError());
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->pin_i.p, c->frame_outi.p, sizeof(int) * ((size_t)n + 8),
                                     cudaMemcpyDeviceToHost, c->stream));
    if (g3)
      SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->pin_d.p, c->frame_outd.p + 2 * (size_t)n, sizeof(double) * 3 * (size_t)n,
                                       cudaMemcpyDeviceToHost, c->stream));
    SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ri = c->pin_i.p;
    rg3 = c->pin_d.p;
  }
  const FrameScalars* sc = reinterpret_cast<const FrameScalars*>(ri);
  std::memcpy(idx, ri + 8, sizeof(int) * n);
  if (g3) std::memcpy(g3, rg3, sizeof(double) * 3 * (size_t)n);
  *cci = (uint32_t)sc->current_cone_index;
  if (n_reobserved) *n_reobserved = sc->n_reobserved;
  if (send_cone_data) *send_cone_data = sc->send_cone_data;
  return sc->n_reobserved;
} SLAM_ABI_CATCH(c)

// Whole drives per replica: see drive_replicas_kernel.  Host buffers in and out; kernel_ms (optional) = device time
// of the drive kernel alone (events on the context's stream).
int slam_b200_drive_replicas(slam_b200_ctx* c, int n_replicas, int n_frames, int nmax, const double* frames4,
                             const int32_t* ncols, const double* poses3, double thr, double mapThr, int cap,
                             int32_t* records, double* map_x, double* map_y, int32_t* map_type, int32_t* map_n,
                             int32_t* closed_at, double* kernel_ms) try {
  NvtxRange nvtx_range("slam_b200/drive_replicas");
  if (!c || n_replicas < 0 || n_frames < 0 || nmax < 1 || cap < 2 || !frames4 || !ncols || !poses3 || !records ||
      !map_n || !closed_at)
    return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  if (n_replicas == 0 || n_frames == 0) return 0;
  const size_t R = (size_t)n_replicas, F = (size_t)n_frames, si = drive_in_stride(nmax), sr = drive_rec_stride(nmax);
  // pack: cone columns, pose, cos / sin of the heading from the host libm (like upload_frame)
  std::vector<double> packed(R * F * si, 0.0);
  for (size_t q = 0; q < R * F; q++) {
    const int n = ncols[q];
    if (n < 0) return SLAM_B200_E_ARG;
    const int m = std::min(n, nmax);
    double* o = packed.data() + q * si;
    std::memcpy(o, frames4 + q * 4 * (size_t)nmax, sizeof(double) * 4 * (size_t)m);
    const double* p = poses3 + 3 * q;
    o[4 * (size_t)m] = p[0]; o[4 * (size_t)m + 1] = p[1]; o[4 * (size_t)m + 2] = p[2];
    o[4 * (size_t)m + 3] = std::cos(p[2]);
    o[4 * (size_t)m + 4] = std::sin(p[2]);
  }
  DevBuf<double> d_frames, d_mx, d_my, d_scratch;
  DevBuf<int> d_ncols, d_mt, d_rec, d_mn, d_closed;
  struct Release {
    DevBuf<double>&a, &b, &c2, &d; DevBuf<int>&e, &f, &g, &h, &i;
    ~Release() { a.release(); b.release(); c2.release(); d.release(); e.release(); f.release(); g.release(); h.release(); i.release(); }
  } rel{d_frames, d_mx, d_my, d_scratch, d_ncols, d_mt, d_rec, d_mn, d_closed};
  SLAM_CUDA_TRY(c, d_frames.exact(R * F * si));
  SLAM_CUDA_TRY(c, d_ncols.exact(R * F));
  SLAM_CUDA_TRY(c, d_mx.exact(R * (size_t)cap));
  SLAM_CUDA_TRY(c, d_my.exact(R * (size_t)cap));
  SLAM_CUDA_TRY(c, d_mt.exact(R * (size_t)cap));
  SLAM_CUDA_TRY(c, d_scratch.exact(R * (8 * (size_t)nmax + 8)));
  SLAM_CUDA_TRY(c, d_rec.exact(R * F * sr));
  SLAM_CUDA_TRY(c, d_mn.exact(R));
  SLAM_CUDA_TRY(c, d_closed.exact(R));
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(d_frames.p, packed.data(), sizeof(double) * R * F * si, cudaMemcpyHostToDevice, c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(d_ncols.p, ncols, sizeof(int) * R * F, cudaMemcpyHostToDevice, c->stream));
  SLAM_CUDA_TRY(c, cudaMemsetAsync(d_mx.p, 0, sizeof(double) * R * (size_t)cap, c->stream));
  SLAM_CUDA_TRY(c, cudaMemsetAsync(d_my.p, 0, sizeof(double) * R * (size_t)cap, c->stream));
  SLAM_CUDA_TRY(c, cudaMemsetAsync(d_mt.p, 0, sizeof(int) * R * (size_t)cap, c->stream));
  cudaEvent_t e0, e1;
  SLAM_CUDA_TRY(c, cudaEventCreate(&e0));
  SLAM_CUDA_TRY(c, cudaEventCreate(&e1));
  cudaEventRecord(e0, c->stream);
  drive_replicas_kernel<<<n_replicas, FRAME_THREADS, 0, c->stream>>>(d_frames.p, d_ncols.p, n_frames, nmax, thr, mapThr,
                                                                     d_mx.p, d_my.p, d_mt.p, cap, d_scratch.p, d_rec.p,
                                                                     d_mn.p, d_closed.p);
  c->launches++;
  cudaEventRecord(e1, c->stream);
  cudaError_t le = cudaGetLastError();
  if (le == cudaSuccess) le = cudaMemcpyAsync(records, d_rec.p, sizeof(int) * R * F * sr, cudaMemcpyDeviceToHost, c->stream);
  if (le == cudaSuccess && map_x) le = cudaMemcpyAsync(map_x, d_mx.p, sizeof(double) * R * (size_t)cap, cudaMemcpyDeviceToHost, c->stream);
  if (le == cudaSuccess && map_y) le = cudaMemcpyAsync(map_y, d_my.p, sizeof(double) * R * (size_t)cap, cudaMemcpyDeviceToHost, c->stream);
  if (le == cudaSuccess && map_type) le = cudaMemcpyAsync(map_type, d_mt.p, sizeof(int) * R * (size_t)cap, cudaMemcpyDeviceToHost, c->stream);
  if (le == cudaSuccess) le = cudaMemcpyAsync(map_n, d_mn.p, sizeof(int) * R, cudaMemcpyDeviceToHost, c->stream);
  if (le == cudaSuccess) le = cudaMemcpyAsync(closed_at, d_closed.p, sizeof(int) * R, cudaMemcpyDeviceToHost, c->stream);
  if (le == cudaSuccess) le = cudaStreamSynchronize(c->stream);
  float ms = 0;
  if (le == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  SLAM_CUDA_TRY(c, le);
  if (kernel_ms) *kernel_ms = ms;
  return n_replicas;
} SLAM_ABI_CATCH(c)

int slam_b200_map_build_grid(slam_b200_ctx* c, double cell) try {
  NvtxRange nvtx_range("slam_b200/map_build_grid");
  if (!c || !(cell > 0)) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  int M = c->map_n;
  c->grid_map_version = 0;
  if (M == 0) {
    c->grid_nx = c->grid_ny = 0;
    c->grid_cell = cell;
    c->grid_map_version = c->map_version;
    return 0;
  }
  SLAM_CUDA_TRY(c, c->grid_bbox.exact(4));
  double init[4] = {DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX}, bb[4];
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->grid_bbox.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  bbox_kernel<<<std::min((M + 255) / 256, 4 * c->num_sms), 256, 0, c->stream>>>(c->map_x.p, c->map_y.p, M, c->grid_bbox.p);
  c->launches++;
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(bb, c->grid_bbox.p, sizeof(bb), cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if (!(bb[0] <= bb[2])) { bb[0] = bb[1] = 0; bb[2] = bb[3] = 0; }  // no finite cone
  // query half-width h = cell * (1 + 1e-9) (covers the rounding of the exact gate, which admits
  // |dx| up to thr * (1 + a few ulp)); cell width = 2 h (1 + 1e-9) so [g-h, g+h] spans <= 2 cells
  // even after the rounding of (v - x0) * inv (relative 2^-52 on <= 1e6 cells)
  GridParams gp;
  gp.h = cell * (1.0 + 1e-9);
  double width = 2.0 * gp.h * (1.0 + 1e-9);
  gp.x0 = bb[0];
  gp.y0 = bb[1];
  gp.inv = 1.0 / width;
  double nxd = std::floor((bb[2] - bb[0]) * gp.inv) + 1, nyd = std::floor((bb[3] - bb[1]) * gp.inv) + 1;
  if (nxd * nyd > 2.0e8) {  // keep the cell table bounded (800 MB of int); coarser cells stay exact
    double f = std::sqrt(nxd * nyd / 2.0e8);
    width *= f;
    gp.inv = 1.0 / width;
    nxd = std::floor((bb[2] - bb[0]) * gp.inv) + 1;
    nyd = std::floor((bb[3] - bb[1]) * gp.inv) + 1;
  }
  gp.nx = (int)nxd;
  gp.ny = (int)nyd;
  size_t ncell = (size_t)gp.nx * gp.ny;
  SLAM_CUDA_TRY(c, c->grid_cell_start.exact(ncell + 1));
  SLAM_CUDA_TRY(c, c->grid_cursor.exact(ncell + 1));
  SLAM_CUDA_TRY(c, c->grid_rec.exact(M));
  SLAM_CUDA_TRY(c, cudaMemsetAsync(c->grid_cursor.p, 0, sizeof(int) * (ncell + 1), c->stream));
  grid_count_kernel<<<(M + 255) / 256, 256, 0, c->stream>>>(c->map_x.p, c->map_y.p, M, gp, c->grid_cursor.p);
  c->launches++;
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, c->grid_cursor.p, c->grid_cell_start.p, (int)(ncell + 1), c->stream);
  SLAM_CUDA_TRY(c, c->grid_tmp.exact(tmp_bytes));
  SLAM_CUDA_TRY(c, cub::DeviceScan::ExclusiveSum(c->grid_tmp.p, tmp_bytes, c->grid_cursor.p, c->grid_cell_start.p,
                                                  (int)(ncell + 1), c->stream));
  c->launches++;
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->grid_cursor.p, c->grid_cell_start.p, sizeof(int) * ncell, cudaMemcpyDeviceToDevice, c->stream));
  grid_fill_kernel<<<(M + 255) / 256, 256, 0, c->stream>>>(c->map_x.p, c->map_y.p, c->map_type.p, M, gp,
                                                          c->grid_cursor.p, c->grid_rec.p);
  c->launches++;
  SLAM_CUDA_TRY(c, cudaGetLastError());
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->grid_cell = cell;
  c->grid_x0 = gp.x0;
  c->grid_y0 = gp.y0;
  c->grid_nx = gp.nx;
  c->grid_ny = gp.ny;
  c->grid_inv = gp.inv;
  c->grid_h = gp.h;
  c->grid_map_version = c->map_version;
  return (int)std::min<size_t>(ncell, 0x7fffffff);
} SLAM_ABI_CATCH(c)

int slam_b200_assoc_bulk_dev(slam_b200_ctx* c, const double* cones_dev, int n, const double pose[3],
                             double thr, int gate, int algo, int32_t* idx_dev) try {
  NvtxRange nvtx_range("slam_b200/assoc_bulk");
  if (!c || !pose || n < 0 || (n > 0 && (!cones_dev || !idx_dev))) return SLAM_B200_E_ARG;
  if (gate != SLAM_B200_GATE_MAPPING && gate != SLAM_B200_GATE_LOCALIZER) return SLAM_B200_E_ARG;
  if (algo != SLAM_B200_ALGO_BRUTE && algo != SLAM_B200_ALGO_GRID && algo != SLAM_B200_ALGO_GRID_PIPELINED)
    return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  if (n == 0) return 0;
  PoseTrig pt{pose[0], pose[1], std::cos(pose[2]), std::sin(pose[2])};
  double thr2x = sqrt_gate_threshold(thr);
  int blocks = (n + BULK_THREADS - 1) / BULK_THREADS;
  if (c->map_n == 0 || !(thr > 0)) {
    SLAM_CUDA_TRY(c, cudaMemsetAsync(idx_dev, 0xff, sizeof(int) * (size_t)n, c->stream));
    return 0;
  }
  if (algo == SLAM_B200_ALGO_BRUTE) {
    if (gate == SLAM_B200_GATE_MAPPING)
      assoc_bulk_brute_kernel<SLAM_B200_GATE_MAPPING><<<blocks, BULK_THREADS, 0, c->stream>>>(
          cones_dev, n, pt, thr2x, c->map_x.p, c->map_y.p, c->map_type.p, c->map_n, idx_dev);
    else
      assoc_bulk_brute_kernel<SLAM_B200_GATE_LOCALIZER><<<blocks, BULK_THREADS, 0, c->stream>>>(
          cones_dev, n, pt, thr2x, c->map_x.p, c->map_y.p, c->map_type.p, c->map_n, idx_dev);
  } else {
    if (c->grid_map_version != c->map_version || !(c->grid_cell >= thr)) {
      int rc = slam_b200_map_build_grid(c, thr);
      if (rc < 0) return rc;
    }
    GridParams gp;
    gp.x0 = c->grid_x0;
    gp.y0 = c->grid_y0;
    gp.nx = c->grid_nx;
    gp.ny = c->grid_ny;
    gp.inv = c->grid_inv;  // the exact values the index was built with
    gp.h = c->grid_h;
    const int* cs = c->grid_cell_start.p;
    const GridRec* rec = c->grid_rec.p;
    if (algo == SLAM_B200_ALGO_GRID_PIPELINED) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(blocks);
      cfg.blockDim = dim3(BULK_THREADS);
      cfg.stream = c->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (gate == SLAM_B200_GATE_MAPPING)
        SLAM_CUDA_TRY(c, cudaLaunchKernelEx(&cfg, assoc_bulk_grid_kernel<SLAM_B200_GATE_MAPPING, true>, cones_dev, n, pt,
                                            thr2x, gp, cs, rec, (int*)idx_dev));
      else
        SLAM_CUDA_TRY(c, cudaLaunchKernelEx(&cfg, assoc_bulk_grid_kernel<SLAM_B200_GATE_LOCALIZER, true>, cones_dev, n, pt,
                                            thr2x, gp, cs, rec, (int*)idx_dev));
    } else if (gate == SLAM_B200_GATE_MAPPING) {
      assoc_bulk_grid_kernel<SLAM_B200_GATE_MAPPING, false><<<blocks, BULK_THREADS, 0, c->stream>>>(
          cones_dev, n, pt, thr2x, gp, cs, rec, idx_dev);
    } else {
      assoc_bulk_grid_kernel<SLAM_B200_GATE_LOCALIZER, false><<<blocks, BULK_THREADS, 0, c->stream>>>(
          cones_dev, n, pt, thr2x, gp, cs, rec, idx_dev);
    }
  }
  c->launches++;
  SLAM_CUDA_TRY(c, cudaGetLastError());
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_assoc_bulk_frames_dev(int n_frames, slam_b200_ctx* const* ctxs, const double* const* cones_dev,
                                    const int* n, const double* poses, double thr, int gate, int algo,
                                    int32_t* const* idx_dev) try {
  if (n_frames < 0 || (n_frames > 0 && (!ctxs || !cones_dev || !n || !poses || !idx_dev))) return SLAM_B200_E_ARG;
  if (algo != SLAM_B200_ALGO_GRID_BATCHED) {
    for (int f = 0; f < n_frames; f++) {
      int rc = slam_b200_assoc_bulk_dev(ctxs[f], cones_dev[f], n[f], poses + 3 * (size_t)f, thr, gate, algo, idx_dev[f]);
      if (rc < 0) return rc;
    }
    return n_frames;
  }
  if (gate != SLAM_B200_GATE_MAPPING && gate != SLAM_B200_GATE_LOCALIZER) return SLAM_B200_E_ARG;
  if (n_frames == 0) return 0;
  slam_b200_ctx* c0 = ctxs[0];
  if (!c0) return SLAM_B200_E_ARG;
  if (ctx_set_device(c0)) return SLAM_B200_E_CUDA;
  const double thr2x = sqrt_gate_threshold(thr);
  for (int f0 = 0; f0 < n_frames; f0 += BATCH_FRAMES) {
    BatchParams bp;
    std::memset(&bp, 0, sizeof(bp));
    int nb = 0, nmax = 0;
    for (int f = f0; f < n_frames && f < f0 + BATCH_FRAMES; f++) {
      slam_b200_ctx* c = ctxs[f];
      if (!c || c->device != c0->device || n[f] < 0 || (n[f] > 0 && (!cones_dev[f] || !idx_dev[f]))) return SLAM_B200_E_ARG;
      if (n[f] == 0) continue;
      if (c->map_n == 0 || !(thr > 0)) {  // nothing can match
        SLAM_CUDA_TRY(c0, cudaMemsetAsync(idx_dev[f], 0xff, sizeof(int) * (size_t)n[f], c0->stream));
        continue;
      }
      if (c->grid_map_version != c->map_version || !(c->grid_cell >= thr)) {
        int rc = slam_b200_map_build_grid(c, thr);  // on c's own stream, synchronised before it returns
        if (rc < 0) return rc;
      }
      BatchFrame& F = bp.f[nb++];
      const double* pose = poses + 3 * (size_t)f;
      F.cones = cones_dev[f];
      F.cell_start = c->grid_cell_start.p;
      F.rec = c->grid_rec.p;
      F.idx = idx_dev[f];
      F.pt = PoseTrig{pose[0], pose[1], std::cos(pose[2]), std::sin(pose[2])};
      F.gp.x0 = c->grid_x0; F.gp.y0 = c->grid_y0; F.gp.nx = c->grid_nx; F.gp.ny = c->grid_ny;
      F.gp.inv = c->grid_inv; F.gp.h = c->grid_h;
      F.n = n[f];
      nmax = std::max(nmax, n[f]);
    }
    if (nb == 0) continue;
    dim3 grid((nmax + BULK_THREADS - 1) / BULK_THREADS, nb);
    if (gate == SLAM_B200_GATE_MAPPING)
      assoc_bulk_grid_batched_kernel<SLAM_B200_GATE_MAPPING><<<grid, BULK_THREADS, 0, c0->stream>>>(bp, thr2x);
    else
      assoc_bulk_grid_batched_kernel<SLAM_B200_GATE_LOCALIZER><<<grid, BULK_THREADS, 0, c0->stream>>>(bp, thr2x);
    c0->launches++;
    SLAM_CUDA_TRY(c0, cudaGetLastError());
  }
  return n_frames;
} catch (...) { return SLAM_B200_E_STATE; }

int slam_b200_assoc_bulk(slam_b200_ctx* c, const double* cones, int n, const double pose[3], double thr,
                         int gate, int algo, int32_t* idx) try {
  if (!c || !pose || n < 0 || (n > 0 && (!cones || !idx))) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  if (n == 0) return 0;
  SLAM_CUDA_TRY(c, c->frame_in.exact(4 * (size_t)n + 4));
  SLAM_CUDA_TRY(c, c->frame_outi.exact(2 * (size_t)n + 8));
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->frame_in.p, cones, sizeof(double) * 4 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  int rc = slam_b200_assoc_bulk_dev(c, c->frame_in.p, n, pose, thr, gate, algo, c->frame_outi.p);
  if (rc < 0) return rc;
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(idx, c->frame_outi.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
} SLAM_ABI_CATCH(c)

}  // extern "C"

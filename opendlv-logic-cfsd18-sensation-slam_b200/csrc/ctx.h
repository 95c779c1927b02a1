// ctx.h -- context object behind the opaque slam_b200_ctx handle (include/slam_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: a pointer check per call when no tool is attached

#include "../../include/slam_b200.h"
#include "symbolic.h"

// Named range for profilers (Nsight Systems timelines, `ncu --nvtx --nvtx-include "slam_b200/..."`): one per
// ABI-level phase.  Host-side only -- kernels replayed from the captured CUDA graph are attributed to the
// range of the call that launches the graph.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

#define SLAM_CUDA_TRY(ctx, expr)                                                      \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      (ctx)->fail(std::string(#expr) + ": " + cudaGetErrorString(_e));                \
      return SLAM_B200_E_CUDA;                                                        \
    }                                                                                 \
  } while (0)

// No C++ exception crosses the C ABI (include/slam_b200.h, "never throw across the ABI"): every entry
// point that can allocate host memory is a function-try-block ending in SLAM_ABI_CATCH(ctx), which
// turns std::bad_alloc into SLAM_B200_E_NOMEM and anything else into SLAM_B200_E_STATE, with the
// message left on the context for slam_b200_last_error().
struct slam_b200_ctx;
int slam_abi_caught(slam_b200_ctx* c) noexcept;  // capi.cu; call only from inside a catch block
#define SLAM_ABI_CATCH(ctx) catch (...) { return slam_abi_caught(ctx); }

// ---- guard-band mode (SLAM_B200_GUARD=1; debugging and tests only) ------------------------------------
// compute-sanitizer is closed on the pool this was developed on, so the library carries its own check:
// with the variable set, every device array is allocated with GUARD_BYTES of 0xFF in front of and behind
// the payload and the payload itself is filled with 0xFF as well (as fp64 a NaN, as int32 -1), so
//  * an out-of-bounds WRITE within the band is caught by slam_b200_debug_guard_check (bands re-read
//    and compared byte for byte),
//  * an out-of-bounds or uninitialised READ brings a NaN / -1 into the arithmetic and shows up in the
//    parity tests (NaN estimates, index -1 faults),
// while the default build pays one predictable branch per allocation.
constexpr size_t GUARD_BYTES = 4096;
bool guard_mode();                                                  // capi.cu
void guard_register(void* raw, size_t payload_bytes);               // capi.cu
void guard_unregister(void* raw);                                   // capi.cu
inline cudaError_t guarded_malloc(void** p, void** raw, size_t bytes) {
  if (!guard_mode()) {
    cudaError_t e = cudaMalloc(raw, bytes);
    *p = *raw;
    return e;
  }
  const size_t padded = (bytes + 255) & ~(size_t)255;
  cudaError_t e = cudaMalloc(raw, padded + 2 * GUARD_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaMemset(*raw, 0xFF, padded + 2 * GUARD_BYTES);
  if (e != cudaSuccess) { cudaFree(*raw); *raw = nullptr; return e; }
  *p = static_cast<char*>(*raw) + GUARD_BYTES;
  guard_register(*raw, bytes);
  return cudaSuccess;
}
inline void guarded_free(void* raw) {
  if (!raw) return;
  if (guard_mode()) guard_unregister(raw);
  cudaFree(raw);
}

// growable device array; grow() keeps the first `keep` elements
template <class T>
struct DevBuf {
  T* p = nullptr;
  void* raw = nullptr;  // what cudaMalloc returned (== p unless guard-band mode is on)
  size_t cap = 0;
  cudaError_t reserve(size_t n, size_t keep, cudaStream_t s) {
    if (n <= cap) return cudaSuccess;
    size_t ncap = cap ? cap : 256;
    while (ncap < n) ncap *= 2;
    void *q = nullptr, *qraw = nullptr;
    cudaError_t e = guarded_malloc(&q, &qraw, ncap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p) {
      e = cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) {
        guarded_free(qraw);
        return e;
      }
    }
    guarded_free(raw);
    p = static_cast<T*>(q);
    raw = qraw;
    cap = ncap;
    return cudaSuccess;
  }
  cudaError_t exact(size_t n) {  // (re)allocate exactly, contents dropped
    if (n <= cap) return cudaSuccess;
    guarded_free(raw);
    p = nullptr;
    raw = nullptr;
    cap = 0;
    void* q = nullptr;
    cudaError_t e = guarded_malloc(&q, &raw, (n ? n : 1) * sizeof(T));
    if (e == cudaSuccess) { p = static_cast<T*>(q); cap = n ? n : 1; }
    return e;
  }
  void release() {
    guarded_free(raw);
    p = nullptr;
    raw = nullptr;
    cap = 0;
  }
};

template <class T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    size_t ncap = cap ? cap : 1024;
    while (ncap < n) ncap *= 2;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(&p, ncap * sizeof(T));
    if (e == cudaSuccess) cap = ncap;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// one map cone in the cell-sorted index: exactly one aligned 32-byte DRAM sector
struct alignas(32) GridRec {
  double x, y;
  int type, idx, pad0, pad1;
};

// vertex id -> (local index << 1) | is_landmark.  The reference numbers cones from 0 and poses from 1000 upwards
// (slam.cpp:434,527), so the ids of a graph span a range about as large as their count: a flat table then, a hash map
// only when the ids are sparse (10,300 hash inserts were a third of graph_load on the 10-lap graph).
struct IdIndex {
  std::vector<int> flat;                 // -1 = absent
  long lo = 0;                           // id of flat[0]
  size_t n = 0;                          // ids stored
  bool hashed = false;
  std::unordered_map<int, int> map;
  static bool dense(long span, size_t count) { return span <= 8L * (long)count + 4096; }
  void clear() { flat.clear(); map.clear(); lo = 0; n = 0; hashed = false; }
  int get(int id) const {
    if (hashed) { auto it = map.find(id); return it == map.end() ? -1 : it->second; }
    const long k = (long)id - lo;
    return (k < 0 || k >= (long)flat.size()) ? -1 : flat[(size_t)k];
  }
  bool put(int id, int v) {              // false: the id is there already
    if (!hashed) {
      if (flat.empty()) { lo = id; flat.assign(64, -1); }
      long k = (long)id - lo;
      if (k < 0 || k >= (long)flat.size()) {
        const long nlo = std::min<long>(lo, id), nhi = std::max<long>(lo + (long)flat.size() - 1, id);
        if (dense(nhi - nlo + 1, n + 1)) {
          // grow geometrically on the side the id fell off
          const long span = nhi - nlo + 1, cap = std::max<long>(span, 2 * (long)flat.size());
          const long flo = id < lo ? nhi - cap + 1 : nlo;
          std::vector<int> g((size_t)cap, -1);
          std::copy(flat.begin(), flat.end(), g.begin() + (lo - flo));
          flat.swap(g);
          lo = flo;
          k = (long)id - lo;
        } else {
          hashed = true;
          map.reserve(2 * n + 16);
          for (size_t q = 0; q < flat.size(); q++)
            if (flat[q] >= 0) map.emplace((int)(lo + (long)q), flat[q]);
          flat.clear();
        }
      }
      if (!hashed) {
        if (flat[(size_t)k] >= 0) return false;
        flat[(size_t)k] = v;
        n++;
        return true;
      }
    }
    if (!map.emplace(id, v).second) return false;
    n++;
    return true;
  }
};

// ---- host-side graph (insertion order preserved; the reference's g2o graph owns the same data) --
struct HostGraph {
  // vertices
  std::vector<int> pose_id, lm_id;
  std::vector<double> pose_est;  // 3 per pose
  std::vector<double> lm_est;    // 2 per landmark
  std::vector<char> pose_fixed, lm_fixed;
  IdIndex id2v;  // id -> (local index << 1) | is_landmark
  // pose-pose edges (EdgeSE2)
  std::vector<int> eo_i, eo_j;     // local pose indices
  std::vector<double> eo_z;        // 3 per edge
  std::vector<double> eo_info;     // 6 per edge: (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
  // pose-landmark edges (EdgeSE2PointXY)
  std::vector<int> el_p, el_l;     // local pose / landmark indices
  std::vector<double> el_z;        // 2 per edge
  std::vector<double> el_info;     // 3 per edge: (0,0) (0,1) (1,1)
  uint64_t structure_version = 1;  // bumped by anything that changes topology / gauge
  uint64_t values_version = 1;     // bumped by anything that changes numbers on the host
  int P() const { return (int)pose_id.size(); }
  int L() const { return (int)lm_id.size(); }
  int Eo() const { return (int)eo_i.size(); }
  int El() const { return (int)el_p.size(); }
};

// ---- device-side system of one topology (see graph.cu / solver.cu) ------------------------------
struct DeviceSystem;  // defined in graph_dev.h

struct slam_b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  long launches = 0;
  int num_sms = 148;
  int max_smem_optin = 0;
  bool solver_attrs_set = false;  // solver.cu: cudaFuncSetAttribute done for this context's device
  // solver.cu: side streams + events that let the size classes of one tree level run side by side (fork / join
  // around the context's stream; inside a stream capture they become parallel branches of the CUDA graph)
  static constexpr int kAuxStreams = 3;
  cudaStream_t aux_stream[kAuxStreams] = {nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> fork_events;

  // ---- cone map ----
  DevBuf<double> map_x, map_y;
  DevBuf<int> map_type;
  int map_n = 0;
  uint64_t map_version = 1;
  // uniform grid index over the map (assoc.cu)
  DevBuf<int> grid_cell_start;   // ncell + 1
  DevBuf<int> grid_cursor;       // ncell
  DevBuf<GridRec> grid_rec;      // cones sorted by cell, one 32-byte record each
  DevBuf<double> grid_bbox;      // 4 doubles: minx, miny, maxx, maxy
  DevBuf<char> grid_tmp;         // scan workspace
  double grid_cell = 0, grid_x0 = 0, grid_y0 = 0, grid_inv = 0, grid_h = 0;
  int grid_nx = 0, grid_ny = 0;
  uint64_t grid_map_version = 0;

  // pinned host mirror of the cone map, refreshed by an asynchronous copy behind an event (assoc.cu:
  // slam_b200_map_update_from_graph / slam_b200_map_mirror): readers wait for that copy only, never for the stream
  PinBuf<double> mirror_xy;      // x[cap] | y[cap]
  PinBuf<int> mirror_type;
  size_t mirror_cap = 0;         // cones the two buffers are laid out for
  int mirror_n = 0;              // cones of the last refresh
  uint64_t mirror_version = 0;   // map_version of the last refresh
  cudaEvent_t mirror_event = nullptr;
  DevBuf<int> lm_of_map;         // landmark (local index) behind every map cone, -1 = none

  // ---- frame staging ----
  DevBuf<double> frame_in;       // 4n + 3
  DevBuf<double> frame_outd;     // 5n doubles: z2, g3
  DevBuf<int> frame_outi;        // 2n + 8 ints
  PinBuf<double> pin_d;
  PinBuf<int> pin_i;
  PinBuf<char> pin_stage;        // structure uploads
  // per-frame mailbox (assoc.cu): mapped pinned memory the frame kernels read their input from and publish their
  // records + a completion word into, so a frame is one launch and no copy / stream synchronisation
  char* mbox_h = nullptr;        // host address
  char* mbox_d = nullptr;        // device address of the same memory
  size_t mbox_cap = 0;           // columns it is sized for
  unsigned mbox_seq = 0;         // completion word the next frame publishes

  // ---- graph ----
  HostGraph g;
  DeviceSystem* sys = nullptr;
  bool assembly_only = false;  // prepare without the symbolic phase (assembly measurements)
  bool batch_ordering = false; // structure is being built for a replica batch (fill-minimising ordering)

  void fail(const std::string& m) { err = m; }
};

int ctx_set_device(slam_b200_ctx* c);
void graph_release(slam_b200_ctx* c);  // graph.cu
// graph.cu: device estimates of replica 0 (x[P] | y[P] | theta[P] | lx[L] | ly[L]) if they are current, i.e. the
// host graph has not changed since the device last wrote them back; false otherwise
bool graph_device_estimates(slam_b200_ctx* c, double** est, int* P, int* L);

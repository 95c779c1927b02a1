// tsan_slam_stub_backend.cpp -- TEST INFRASTRUCTURE ONLY, linked into one ThreadSanitizer binary
// (tests/test_frame_assembler.py::test_slam_lock_discipline_under_tsan).  NOT a CPU fallback and not a
// restatement of anything: the sixteen C-ABI entry points the host Slam class calls, answering with
// canned records (every column matches or creates the next cone, the loop "closes" at frame 20, the
// optimiser "runs" 10 iterations) so that a CPU-only box can drive csrc/host/slam.cpp through its
// mapping, loop-closing and localiser branches from two threads and let TSan watch the mutexes.
#include <cstdint>
#include <cstring>

#include "../include/slam_b200.h"

struct slam_b200_ctx {
  int map_n = 0;
  int frames = 0;
};

extern "C" {
int slam_b200_create(int, void*, slam_b200_ctx** out) { *out = new slam_b200_ctx(); return 0; }
int slam_b200_destroy(slam_b200_ctx* c) { delete c; return 0; }
int slam_b200_warmup(slam_b200_ctx*, int, int) { return 0; }
const char* slam_b200_last_error(const slam_b200_ctx*) { return "stub"; }
int slam_b200_graph_clear(slam_b200_ctx*) { return 0; }
int slam_b200_map_clear(slam_b200_ctx* c) { c->map_n = 0; return 0; }
int slam_b200_graph_add_pose(slam_b200_ctx*, int, double, double, double) { return 0; }
int slam_b200_graph_add_odometry(slam_b200_ctx*, int, int, const double*, const double*) { return 0; }
int slam_b200_graph_add_landmark(slam_b200_ctx*, int, double, double) { return 0; }
int slam_b200_graph_add_edge_se2_xy(slam_b200_ctx*, int, int, const double*, const double*) { return 0; }
int slam_b200_graph_set_fixed(slam_b200_ctx*, int, int) { return 0; }
int slam_b200_graph_optimize(slam_b200_ctx*, int iters, double* chi2) {
  for (int k = 0; k < iters; k++) chi2[k] = 1.0 / (k + 1);
  return iters;
}
int slam_b200_graph_get_vertex(slam_b200_ctx*, int id, double out[3]) {
  out[0] = id * 0.5; out[1] = -id * 0.25; out[2] = 0.1;
  return id < 1000 ? 2 : 3;
}
int slam_b200_map_write_xy(slam_b200_ctx*, int, int n, const double*, const double*) { return n; }
int slam_b200_map_update_from_graph(slam_b200_ctx*) { return 0; }
int slam_b200_map_mirror(slam_b200_ctx*, const double** x, const double** y, const int32_t** t) {
  static const double zero = 0;
  static const int32_t zt = 0;
  if (x) *x = &zero;
  if (y) *y = &zero;
  if (t) *t = &zt;
  return 0;
}
int slam_b200_cones_to_global(slam_b200_ctx*, const double*, int n, const double*, double* g3, double* l3) {
  if (g3) std::memset(g3, 0, sizeof(double) * 3 * (size_t)n);
  if (l3) std::memset(l3, 0, sizeof(double) * 3 * (size_t)n);
  return 0;
}
int slam_b200_assoc_map_frame(slam_b200_ctx* c, const double*, int n, const double*, double, double, uint32_t* cci,
                              int32_t* lc, int32_t* idx, int32_t* status, double* z2, double* g3, int32_t* first,
                              int32_t* lcObs) {
  *first = 0;
  *lcObs = -1;
  if (n == 0) return c->map_n;
  if (c->map_n == 0) { *first = 1; c->map_n = 1; }
  for (int i = 0; i < n; i++) {
    z2[2 * i] = i; z2[2 * i + 1] = -i;
    g3[3 * i] = c->frames + i; g3[3 * i + 1] = i; g3[3 * i + 2] = 1 + (i & 1);
    if (i < c->map_n && (i + c->frames) % 3) { idx[i] = i; status[i] = SLAM_B200_ASSOC_MATCHED; }
    else { idx[i] = c->map_n++; status[i] = SLAM_B200_ASSOC_NEW; }
  }
  *cci = (uint32_t)(c->frames % c->map_n);
  if (++c->frames >= 20) { *lc = 1; *lcObs = n - 1; }
  return c->map_n;
}
int slam_b200_assoc_localize_frame(slam_b200_ctx* c, const double*, int n, const double*, double, uint32_t* cci,
                                   int32_t* idx, double*, int32_t* reobs, int32_t* send) {
  for (int i = 0; i < n; i++) idx[i] = (i & 1) ? -1 : (i + c->frames) % c->map_n;
  c->frames++;
  *cci = (uint32_t)(c->frames % c->map_n);
  *reobs = (n + 1) / 2;
  *send = 1;
  return *reobs;
}
}

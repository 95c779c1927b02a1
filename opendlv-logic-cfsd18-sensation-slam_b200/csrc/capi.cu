// capi.cu -- context management and the graph half of the C ABI (include/slam_b200.h).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <map>
#include <memory>
#include <mutex>
#include <new>

#include "graph_dev.h"
#include "tileplan.h"

// ---- guard-band mode (ctx.h): registry of live device arrays --------------------------------------
namespace {
std::mutex g_guard_mu;
std::unordered_map<void*, size_t>& guard_tab() {
  static std::unordered_map<void*, size_t> t;
  return t;
}
}  // namespace
bool guard_mode() {
  static const bool on = [] { const char* e = getenv("SLAM_B200_GUARD"); return e && atoi(e) != 0; }();
  return on;
}
void guard_register(void* raw, size_t payload_bytes) {
  std::lock_guard<std::mutex> lk(g_guard_mu);
  guard_tab()[raw] = payload_bytes;
}
void guard_unregister(void* raw) {
  std::lock_guard<std::mutex> lk(g_guard_mu);
  guard_tab().erase(raw);
}

int ctx_set_device(slam_b200_ctx* c) {
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) {
    c->fail(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return SLAM_B200_E_CUDA;
  }
  return 0;
}

namespace {

// g2o stuff/misc.h normalize_theta and types/slam2d/se2.h inverse()/operator* on the host: only
// used to form the odometry measurement when a pose is added (Slam::addOdometryMeasurement,
// slam.cpp:452-455) -- graph construction, not the hot path.
double normalize_theta_host(double theta) {
  if (theta >= -M_PI && theta < M_PI) return theta;
  double multiplier = std::floor(theta / (2 * M_PI));
  theta = theta - multiplier * 2 * M_PI;
  if (theta >= M_PI) theta -= 2 * M_PI;
  if (theta < -M_PI) theta += 2 * M_PI;
  return theta;
}
void se2_between(const double* prev, const double* cur, double* z) {
  double ith = normalize_theta_host(-prev[2]);
  double s = std::sin(ith), c = std::cos(ith);
  double itx = c * (-1 * prev[0]) - s * (-1 * prev[1]);
  double ity = s * (-1 * prev[0]) + c * (-1 * prev[1]);
  z[0] = itx + (c * cur[0] - s * cur[1]);
  z[1] = ity + (s * cur[0] + c * cur[1]);
  z[2] = normalize_theta_host(ith + cur[2]);
}

int lookup(slam_b200_ctx* c, int id, bool want_lm, int* local) {
  const int v = c->g.id2v.get(id);
  if (v < 0) return SLAM_B200_E_ARG;
  if (((v & 1) != 0) != want_lm) return SLAM_B200_E_ARG;
  *local = v >> 1;
  return 0;
}

int refresh_host_estimates(slam_b200_ctx* c) {  // device replica 0 -> host graph
  DeviceSystem& D = *c->sys;
  HostGraph& g = c->g;
  size_t ne = (size_t)D.estStride;
  if (!ne) return 0;
  SLAM_CUDA_TRY(c, c->pin_d.reserve(ne));
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->pin_d.p, D.est.p, sizeof(double) * ne, cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  const double* e = c->pin_d.p;
  const int P = D.P, L = D.L;
  for (int p = 0; p < P; p++) {
    g.pose_est[3 * (size_t)p] = e[p];
    g.pose_est[3 * (size_t)p + 1] = e[P + p];
    g.pose_est[3 * (size_t)p + 2] = e[2 * (size_t)P + p];
  }
  for (int l = 0; l < L; l++) {
    g.lm_est[2 * (size_t)l] = e[3 * (size_t)P + l];
    g.lm_est[2 * (size_t)l + 1] = e[3 * (size_t)P + L + l];
  }
  return 0;
}

}  // namespace

// see SLAM_ABI_CATCH in ctx.h
int slam_abi_caught(slam_b200_ctx* c) noexcept {
  int rc = SLAM_B200_E_STATE;
  const char* what = "unknown C++ exception";
  try {
    throw;
  } catch (const std::bad_alloc&) {
    rc = SLAM_B200_E_NOMEM;
    what = "out of host memory";
  } catch (const std::exception& e) {
    what = e.what();
  } catch (...) {
  }
  if (c) {
    try {
      c->fail(std::string("exception at the C ABI: ") + what);
    } catch (...) {
    }
  }
  return rc;
}

extern "C" {

int slam_b200_version(void) { return 100; }

int slam_b200_create(int device, void* stream, slam_b200_ctx** out) {
  if (!out) return SLAM_B200_E_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SLAM_B200_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return SLAM_B200_E_CUDA;
  slam_b200_ctx* c = new (std::nothrow) slam_b200_ctx();
  if (!c) return SLAM_B200_E_NOMEM;
  c->device = device;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return SLAM_B200_E_CUDA;
    }
    c->own_stream = true;
  }
  cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  *out = c;
  return 0;
}

int slam_b200_destroy(slam_b200_ctx* c) {
  if (!c) return SLAM_B200_E_ARG;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  graph_release(c);
  c->map_x.release(); c->map_y.release(); c->map_type.release();
  c->grid_cell_start.release(); c->grid_cursor.release(); c->grid_rec.release();
  c->grid_bbox.release(); c->grid_tmp.release();
  c->frame_in.release(); c->frame_outd.release(); c->frame_outi.release();
  c->pin_d.release(); c->pin_i.release(); c->pin_stage.release();
  c->mirror_xy.release(); c->mirror_type.release(); c->lm_of_map.release();
  if (c->mirror_event) cudaEventDestroy(c->mirror_event);
  c->mirror_event = nullptr;
  if (c->mbox_h) cudaFreeHost(c->mbox_h);
  c->mbox_h = c->mbox_d = nullptr;
  for (cudaStream_t& a : c->aux_stream) if (a) { cudaStreamDestroy(a); a = nullptr; }
  for (cudaEvent_t e : c->fork_events) cudaEventDestroy(e);
  c->fork_events.clear();
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int slam_b200_warmup(slam_b200_ctx* c, int poses_hint, int landmarks_hint) try {
  NvtxRange nvtx_range("slam_b200/warmup");
  if (!c) return SLAM_B200_E_ARG;
  if (c->g.P() || c->g.L() || c->map_n) { c->fail("warmup needs an empty context"); return SLAM_B200_E_STATE; }
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  const int P = std::max(8, poses_hint), L = std::max(8, landmarks_hint) & ~1, half = L / 2;
  const int base = std::max(1000, L);  // pose ids behind the landmark ids, like the reference (slam.hpp:118)
  const double two_pi = 6.283185307179586, R = std::max(4.0, P * 0.45 / two_pi);
  const double info3[9] = {5, 0, 0, 0, 5, 0, 0, 0, 5}, info2[4] = {0.01, 0, 0, 0.01};
  int rc = 0;
  std::vector<double> lx(L), ly(L);
  for (int j = 0; j < L && rc >= 0; j++) {  // cone pairs left and right of the ring
    const double a = two_pi * (j / 2) / half, r = R + ((j & 1) ? 1.5 : -1.5);
    lx[j] = r * std::cos(a);
    ly[j] = r * std::sin(a);
    rc = slam_b200_graph_add_landmark(c, j, lx[j], ly[j]);
  }
  for (int k = 0; k < P && rc >= 0; k++) {
    const double a = two_pi * k / P, pose[3] = {R * std::cos(a), R * std::sin(a), a + 0.5 * 3.141592653589793};
    rc = slam_b200_graph_add_pose(c, base + k, pose[0], pose[1], pose[2] > 3.141592653589793 ? pose[2] - two_pi : pose[2]);
    if (rc >= 0 && k > 0) rc = slam_b200_graph_add_odometry(c, base + k - 1, base + k, pose, info3);
    const int pair0 = (int)((long)k * half / P);
    const double ct = std::cos(pose[2]), st = std::sin(pose[2]);
    for (int q = 0; q < 8 && rc >= 0; q++) {  // the four pairs ahead: eight observations per pose, like a trackdrive frame
      const int j = (2 * (pair0 + 1 + q / 2) + (q & 1)) % L;
      const double dx = lx[j] - pose[0], dy = ly[j] - pose[1], z[2] = {ct * dx + st * dy, -st * dx + ct * dy};
      rc = slam_b200_graph_add_edge_se2_xy(c, base + k, j, z, info2);
    }
  }
  if (rc >= 0) {
    slam_b200_graph_set_fixed(c, base, 1);
    slam_b200_graph_set_fixed(c, base + 1, 1);
    slam_b200_graph_set_fixed(c, 0, 1);
    slam_b200_graph_set_fixed(c, 1, 1);
    double chi2[2];
    rc = slam_b200_graph_optimize(c, 2, chi2);
    if (rc >= -1) rc = 0;  // g2o's -1 / 0 are not errors of the warm-up
  }
  // one mapping frame into the empty map, one localiser frame against it
  if (rc >= 0) {
    const double cones[8] = {10.0, 0.0, 5.0, 1.0, -12.0, 0.0, 6.0, 2.0}, pose[3] = {0, 0, 0};
    uint32_t cci = 0;
    int32_t lc = 0, idx[2], status[2], first = 0, lcobs = -1, reobs = 0, send = 0;
    double z2[4], g3[6];
    rc = slam_b200_assoc_map_frame(c, cones, 2, pose, 1.2, 50.0, &cci, &lc, idx, status, z2, g3, &first, &lcobs);
    if (rc >= 0) rc = slam_b200_assoc_localize_frame(c, cones, 2, pose, 1.2, &cci, idx, g3, &reobs, &send);
    if (rc >= 0) rc = slam_b200_cones_to_global(c, cones, 2, pose, g3, nullptr);
  }
  const std::string keep = c->err;
  slam_b200_graph_clear(c);
  slam_b200_map_clear(c);
  c->err = keep;
  return rc < 0 ? rc : 0;
} SLAM_ABI_CATCH(c)

const char* slam_b200_last_error(const slam_b200_ctx* c) { return c ? c->err.c_str() : "null context"; }

int slam_b200_sync(slam_b200_ctx* c) try {
  if (!c) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
} SLAM_ABI_CATCH(c)

long slam_b200_launch_count(slam_b200_ctx* c) { return c ? c->launches : 0; }

// ------------------------------------------------------------------------------------------------
// graph construction (host mirror; uploaded by prepare)
// ------------------------------------------------------------------------------------------------
int slam_b200_graph_clear(slam_b200_ctx* c) try {
  if (!c) return SLAM_B200_E_ARG;
  uint64_t sv = c->g.structure_version + 1, vv = c->g.values_version + 1;
  c->g = HostGraph();
  c->g.structure_version = sv;
  c->g.values_version = vv;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_add_pose(slam_b200_ctx* c, int id, double x, double y, double theta) try {
  if (!c) return SLAM_B200_E_ARG;
  HostGraph& g = c->g;
  if (!g.id2v.put(id, g.P() << 1)) { c->fail("duplicate vertex id"); return SLAM_B200_E_ARG; }
  g.pose_id.push_back(id);
  g.pose_est.push_back(x); g.pose_est.push_back(y); g.pose_est.push_back(theta);
  g.pose_fixed.push_back(0);
  g.structure_version++;
  g.values_version++;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_add_landmark(slam_b200_ctx* c, int id, double x, double y) try {
  if (!c) return SLAM_B200_E_ARG;
  HostGraph& g = c->g;
  if (!g.id2v.put(id, (g.L() << 1) | 1)) { c->fail("duplicate vertex id"); return SLAM_B200_E_ARG; }
  g.lm_id.push_back(id);
  g.lm_est.push_back(x); g.lm_est.push_back(y);
  g.lm_fixed.push_back(0);
  g.structure_version++;
  g.values_version++;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_add_edge_se2(slam_b200_ctx* c, int id_from, int id_to, const double z[3], const double info[9]) try {
  if (!c || !z || !info) return SLAM_B200_E_ARG;
  int i, j;
  if (lookup(c, id_from, false, &i) || lookup(c, id_to, false, &j) || i == j) {
    c->fail("EdgeSE2: unknown or identical pose ids");
    return SLAM_B200_E_ARG;
  }
  HostGraph& g = c->g;
  g.eo_i.push_back(i);
  g.eo_j.push_back(j);
  for (int k = 0; k < 3; k++) g.eo_z.push_back(z[k]);
  // information matrices are symmetric; the upper triangle is kept
  g.eo_info.push_back(info[0]); g.eo_info.push_back(info[1]); g.eo_info.push_back(info[2]);
  g.eo_info.push_back(info[4]); g.eo_info.push_back(info[5]); g.eo_info.push_back(info[8]);
  g.structure_version++;
  g.values_version++;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_add_odometry(slam_b200_ctx* c, int id_prev, int id_cur, const double pose[3], const double info[9]) try {
  if (!c || !pose || !info) return SLAM_B200_E_ARG;
  int i;
  if (lookup(c, id_prev, false, &i)) { c->fail("odometry: unknown previous pose"); return SLAM_B200_E_ARG; }
  double z[3];
  se2_between(&c->g.pose_est[3 * (size_t)i], pose, z);
  return slam_b200_graph_add_edge_se2(c, id_prev, id_cur, z, info);
} SLAM_ABI_CATCH(c)

int slam_b200_graph_add_edge_se2_xy(slam_b200_ctx* c, int pose_id, int landmark_id, const double z[2], const double info[4]) try {
  if (!c || !z || !info) return SLAM_B200_E_ARG;
  int p, l;
  if (lookup(c, pose_id, false, &p) || lookup(c, landmark_id, true, &l)) {
    c->fail("EdgeSE2PointXY: unknown pose or landmark id");
    return SLAM_B200_E_ARG;
  }
  HostGraph& g = c->g;
  g.el_p.push_back(p);
  g.el_l.push_back(l);
  g.el_z.push_back(z[0]); g.el_z.push_back(z[1]);
  g.el_info.push_back(info[0]); g.el_info.push_back(info[1]); g.el_info.push_back(info[3]);
  g.structure_version++;
  g.values_version++;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_set_fixed(slam_b200_ctx* c, int id, int fixed) try {
  if (!c) return SLAM_B200_E_ARG;
  const int v = c->g.id2v.get(id);
  if (v < 0) { c->fail("setFixed: unknown id"); return SLAM_B200_E_ARG; }
  char f = fixed ? 1 : 0;
  char& cur = (v & 1) ? c->g.lm_fixed[v >> 1] : c->g.pose_fixed[v >> 1];
  if (cur != f) {
    cur = f;
    c->g.structure_version++;
  }
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_load(slam_b200_ctx* c, int n_poses, const int32_t* pose_ids, const double* pose_est3,
                         int n_landmarks, const int32_t* lm_ids, const double* lm_est2, int n_eo,
                         const int32_t* eo_from, const int32_t* eo_to, const double* eo_z3, const double* eo_info9,
                         int n_el, const int32_t* el_pose, const int32_t* el_lm, const double* el_z2,
                         const double* el_info4, int n_fixed, const int32_t* fixed_ids) try {
  NvtxRange nvtx_range("slam_b200/graph_load");
  if (!c || n_poses < 0 || n_landmarks < 0 || n_eo < 0 || n_el < 0 || n_fixed < 0) return SLAM_B200_E_ARG;
  // every array whose count is > 0 must be there: E_ARG, not a crash
  if ((n_poses && (!pose_ids || !pose_est3)) || (n_landmarks && (!lm_ids || !lm_est2)) ||
      (n_eo && (!eo_from || !eo_to || !eo_z3 || !eo_info9)) || (n_el && (!el_pose || !el_lm || !el_z2 || !el_info4)) ||
      (n_fixed && !fixed_ids)) {
    c->fail("graph_load: null array with a non-zero count");
    return SLAM_B200_E_ARG;
  }
  slam_b200_graph_clear(c);
  // any error below leaves an EMPTY graph behind, never a half-built one a later optimise would run on
  struct ClearOnError {
    slam_b200_ctx* c; bool armed = true;
    ~ClearOnError() { if (armed) { std::string keep = c->err; slam_b200_graph_clear(c); c->err = keep; } }
  } guard{c};
  HostGraph& g = c->g;
  g.pose_id.assign(pose_ids, pose_ids + n_poses);
  g.pose_est.assign(pose_est3, pose_est3 + 3 * (size_t)n_poses);
  g.pose_fixed.assign(n_poses, 0);
  g.lm_id.assign(lm_ids, lm_ids + n_landmarks);
  g.lm_est.assign(lm_est2, lm_est2 + 2 * (size_t)n_landmarks);
  g.lm_fixed.assign(n_landmarks, 0);
  {
    // the id index in one go: a flat table over the id range when the ids are dense (IdIndex, ctx.h)
    long lo = INT32_MAX, hi = INT32_MIN;
    for (int p = 0; p < n_poses; p++) { lo = std::min<long>(lo, pose_ids[p]); hi = std::max<long>(hi, pose_ids[p]); }
    for (int l = 0; l < n_landmarks; l++) { lo = std::min<long>(lo, lm_ids[l]); hi = std::max<long>(hi, lm_ids[l]); }
    const size_t count = (size_t)n_poses + n_landmarks;
    IdIndex& ix = g.id2v;
    ix.clear();
    if (count && IdIndex::dense(hi - lo + 1, count)) {
      ix.lo = lo;
      ix.flat.assign((size_t)(hi - lo + 1), -1);
      for (int p = 0; p < n_poses; p++) {
        int& slot = ix.flat[(size_t)(pose_ids[p] - lo)];
        if (slot >= 0) { c->fail("duplicate vertex id"); return SLAM_B200_E_ARG; }
        slot = p << 1;
      }
      for (int l = 0; l < n_landmarks; l++) {
        int& slot = ix.flat[(size_t)(lm_ids[l] - lo)];
        if (slot >= 0) { c->fail("duplicate vertex id"); return SLAM_B200_E_ARG; }
        slot = (l << 1) | 1;
      }
      ix.n = count;
    } else {
      ix.hashed = count > 0;
      ix.map.reserve(count);
      for (int p = 0; p < n_poses; p++)
        if (!ix.map.emplace(pose_ids[p], p << 1).second) { c->fail("duplicate vertex id"); return SLAM_B200_E_ARG; }
      for (int l = 0; l < n_landmarks; l++)
        if (!ix.map.emplace(lm_ids[l], (l << 1) | 1).second) { c->fail("duplicate vertex id"); return SLAM_B200_E_ARG; }
      ix.n = count;
    }
  }
  auto find = [&](int id, bool want_lm, int* local) -> int { return lookup(c, id, want_lm, local); };
  g.eo_i.resize(n_eo); g.eo_j.resize(n_eo);
  g.eo_z.assign(eo_z3, eo_z3 + 3 * (size_t)n_eo);
  g.eo_info.resize(6 * (size_t)n_eo);
  for (int e = 0; e < n_eo; e++) {
    int i, j;
    if (find(eo_from[e], false, &i) || find(eo_to[e], false, &j) || i == j) {
      c->fail("EdgeSE2: unknown or identical pose ids");
      return SLAM_B200_E_ARG;
    }
    g.eo_i[e] = i; g.eo_j[e] = j;
    const double* in = eo_info9 + 9 * (size_t)e;
    double* o = &g.eo_info[6 * (size_t)e];
    o[0] = in[0]; o[1] = in[1]; o[2] = in[2]; o[3] = in[4]; o[4] = in[5]; o[5] = in[8];
  }
  g.el_p.resize(n_el); g.el_l.resize(n_el);
  g.el_z.assign(el_z2, el_z2 + 2 * (size_t)n_el);
  g.el_info.resize(3 * (size_t)n_el);
  for (int e = 0; e < n_el; e++) {
    int p, l;
    if (find(el_pose[e], false, &p) || find(el_lm[e], true, &l)) {
      c->fail("EdgeSE2PointXY: unknown pose or landmark id");
      return SLAM_B200_E_ARG;
    }
    g.el_p[e] = p; g.el_l[e] = l;
    const double* in = el_info4 + 4 * (size_t)e;
    double* o = &g.el_info[3 * (size_t)e];
    o[0] = in[0]; o[1] = in[1]; o[2] = in[3];
  }
  for (int k = 0; k < n_fixed; k++) {
    int rc = slam_b200_graph_set_fixed(c, fixed_ids[k], 1);
    if (rc) return rc;
  }
  g.structure_version++;
  g.values_version++;
  guard.armed = false;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_set_values(slam_b200_ctx* c, const double* pose_est3, const double* lm_est2,
                               const double* eo_z3, const double* el_z2) try {
  if (!c) return SLAM_B200_E_ARG;
  HostGraph& g = c->g;
  if (pose_est3) std::copy(pose_est3, pose_est3 + 3 * (size_t)g.P(), g.pose_est.begin());
  if (lm_est2) std::copy(lm_est2, lm_est2 + 2 * (size_t)g.L(), g.lm_est.begin());
  if (eo_z3) std::copy(eo_z3, eo_z3 + 3 * (size_t)g.Eo(), g.eo_z.begin());
  if (el_z2) std::copy(el_z2, el_z2 + 2 * (size_t)g.El(), g.el_z.begin());
  g.values_version++;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_num_poses(slam_b200_ctx* c) { return c ? c->g.P() : SLAM_B200_E_ARG; }
int slam_b200_graph_num_landmarks(slam_b200_ctx* c) { return c ? c->g.L() : SLAM_B200_E_ARG; }
int slam_b200_graph_num_edges(slam_b200_ctx* c) { return c ? c->g.Eo() + c->g.El() : SLAM_B200_E_ARG; }

int slam_b200_graph_get_vertex(slam_b200_ctx* c, int id, double out[3]) try {
  if (!c || !out) return SLAM_B200_E_ARG;
  const int v = c->g.id2v.get(id);
  if (v < 0) return SLAM_B200_E_ARG;
  int k = v >> 1;
  if (v & 1) {
    out[0] = c->g.lm_est[2 * (size_t)k]; out[1] = c->g.lm_est[2 * (size_t)k + 1]; out[2] = 0;
    return 2;
  }
  out[0] = c->g.pose_est[3 * (size_t)k]; out[1] = c->g.pose_est[3 * (size_t)k + 1]; out[2] = c->g.pose_est[3 * (size_t)k + 2];
  return 3;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_get_estimates(slam_b200_ctx* c, double* pose_est3, double* lm_est2) try {
  if (!c) return SLAM_B200_E_ARG;
  if (pose_est3) std::copy(c->g.pose_est.begin(), c->g.pose_est.end(), pose_est3);
  if (lm_est2) std::copy(c->g.lm_est.begin(), c->g.lm_est.end(), lm_est2);
  return 0;
} SLAM_ABI_CATCH(c)

// ------------------------------------------------------------------------------------------------
// optimisation
// ------------------------------------------------------------------------------------------------
int slam_b200_graph_prepare(slam_b200_ctx* c) try {
  NvtxRange nvtx_range("slam_b200/graph_prepare");
  if (!c) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  int n = graph_build_structure(c);
  if (n < 0) return n;
  DeviceSystem& D = *c->sys;
  if (D.R != 1 || D.values_version != c->g.values_version) {
    int rc = graph_alloc_values(c, 1);
    if (rc) return rc;
    rc = graph_upload_host_values(c);
    if (rc) return rc;
  }
  return n;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_prepare_assembly_only(slam_b200_ctx* c) try {
  if (!c) return SLAM_B200_E_ARG;
  c->assembly_only = true;
  int n = slam_b200_graph_prepare(c);
  c->assembly_only = false;
  return n;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_reset_device(slam_b200_ctx* c) try {
  if (!c || !c->sys) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  int rc = graph_alloc_values(c, 1);
  if (rc) return rc;
  return graph_upload_host_values(c);
} SLAM_ABI_CATCH(c)

int slam_b200_graph_iterate_async(slam_b200_ctx* c, int iters) try {
  NvtxRange nvtx_range("slam_b200/graph_iterate");
  if (!c || !c->sys || iters < 0 || c->sys->assembly_only) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  DeviceSystem& D = *c->sys;
  if (D.n == 0) return 0;
  for (int it = 0; it < iters; it++) {
    int rc = graph_enqueue_iteration(c);
    if (rc) return rc;
  }
  D.iters_enqueued += iters;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_finish(slam_b200_ctx* c, double* chi2, int chi2_cap) try {
  NvtxRange nvtx_range("slam_b200/graph_finish");
  if (!c || !c->sys) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  DeviceSystem& D = *c->sys;
  if (D.n == 0) return -1;
  // chi2 after the last update (g2o verbose: computeActiveErrors after every iteration)
  int rc = graph_enqueue_assemble(c, 0, D.P, true);
  if (rc) return rc;
  int st[2];
  double ch[64];
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(st, D.status.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(ch, D.chi2.p, sizeof(double) * D.chi2_cap, cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  rc = refresh_host_estimates(c);
  if (rc) return rc;
  c->g.values_version++;
  D.values_version = c->g.values_version;  // device and host agree again
  int done = st[1];
  if (chi2)
    for (int k = 0; k < done && k < chi2_cap && k + 1 < D.chi2_cap; k++) chi2[k] = ch[k + 1];
  int failed = st[0];
  // next round of iterations starts a fresh chi2 log / status
  SLAM_CUDA_TRY(c, cudaMemsetAsync(D.status.p, 0, sizeof(int) * 2 * (size_t)D.R, c->stream));
  D.iters_enqueued = 0;
  if (failed) return 0;
  return done;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_optimize(slam_b200_ctx* c, int iters, double* chi2) try {
  NvtxRange nvtx_range("slam_b200/graph_optimize");
  if (!c || iters < 0) return SLAM_B200_E_ARG;
  int n = slam_b200_graph_prepare(c);
  if (n < 0) return n;
  if (n == 0) return -1;  // g2o: nothing to optimise
  int rc = slam_b200_graph_iterate_async(c, iters);
  if (rc) return rc;
  return slam_b200_graph_finish(c, chi2, iters);
} SLAM_ABI_CATCH(c)

int slam_b200_graph_chi2(slam_b200_ctx* c, double* chi2) try {
  if (!c || !chi2) return SLAM_B200_E_ARG;
  int n = slam_b200_graph_prepare(c);
  if (n < 0) return n;
  DeviceSystem& D = *c->sys;
  if (D.n == 0) { *chi2 = 0; return 0; }
  int rc = graph_enqueue_assemble(c, 0, D.P, true);
  if (rc) return rc;
  int st[2];
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(st, D.status.p, sizeof(st), cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  int slot = std::min(st[1], D.chi2_cap - 1);  // the slot the kernel wrote (= iterations done so far)
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(chi2, D.chi2.p + slot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_assemble_async(slam_b200_ctx* c, int p0, int p1) try {
  NvtxRange nvtx_range("slam_b200/graph_assemble");
  if (!c || !c->sys) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  DeviceSystem& D = *c->sys;
  if (p0 < 0 || p1 > D.P || p0 > p1) return SLAM_B200_E_ARG;
  return graph_enqueue_assemble(c, p0, p1, false);
} SLAM_ABI_CATCH(c)

// ---- peer exchange of the landmark part (config 5, one process per GPU) -------------------------
int slam_b200_graph_shard_landmarks(slam_b200_ctx* c, int p0, int p1, int32_t* l0, int32_t* l1) try {
  if (!c || !c->sys || !l0 || !l1) return SLAM_B200_E_STATE;
  DeviceSystem& D = *c->sys;
  if (p0 < 0 || p1 > D.P || p0 > p1) return SLAM_B200_E_ARG;
  int a, b;
  graph_shard_landmarks(D, p0, p1, &a, &b);
  *l0 = a; *l1 = b;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_xchg_create(slam_b200_ctx* c, int world, int rank, int cap, unsigned char handle_out[64]) try {
  if (!c || !c->sys || !handle_out || world < 1 || world > 64 || rank < 0 || rank >= world || cap < 1) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DeviceSystem& D = *c->sys;
  xchg_release(D);
  PeerExchange& X = D.xchg;
  X.world = world; X.rank = rank; X.cap = cap;
  X.bytes = 1024 + sizeof(double) * 2 * 6 * (size_t)cap;
  SLAM_CUDA_TRY(c, cudaMalloc(&X.local, X.bytes));
  SLAM_CUDA_TRY(c, cudaMemset(X.local, 0, X.bytes));
  cudaIpcMemHandle_t h;
  SLAM_CUDA_TRY(c, cudaIpcGetMemHandle(&h, X.local));
  std::memcpy(handle_out, &h, 64);
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_xchg_connect(slam_b200_ctx* c, const unsigned char* handles, const int32_t* ranges) try {
  if (!c || !c->sys || !handles || !ranges || !c->sys->xchg.local) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  PeerExchange& X = c->sys->xchg;
  // validate BEFORE any handle is opened, so an argument error leaves nothing to close
  for (int r = 0; r < X.world; r++)
    if (ranges[2 * r] < 0 || ranges[2 * r + 1] < ranges[2 * r] || ranges[2 * r + 1] - ranges[2 * r] > X.cap ||
        ranges[2 * r + 1] > c->sys->L) return SLAM_B200_E_ARG;
  for (int r = 0; r < (int)X.peers.size(); r++)  // a second connect: drop the first one's mappings
    if (r != X.rank && X.peers[r]) cudaIpcCloseMemHandle(X.peers[r]);
  X.connected = false;
  X.peers.assign(X.world, nullptr);
  for (int r = 0; r < X.world; r++) {
    if (r == X.rank) { X.peers[r] = X.local; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < r; q++)
        if (q != X.rank && X.peers[q]) cudaIpcCloseMemHandle(X.peers[q]);
      X.peers.assign(X.world, nullptr);
      c->fail(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
      return SLAM_B200_E_CUDA;
    }
    X.peers[r] = static_cast<char*>(p);
  }
  X.ranges_host.assign(ranges, ranges + 2 * (size_t)X.world);
  SLAM_CUDA_TRY(c, X.peer_tab.exact(X.world));
  SLAM_CUDA_TRY(c, X.ranges.exact(2 * (size_t)X.world));
  SLAM_CUDA_TRY(c, X.err.exact(1));
  SLAM_CUDA_TRY(c, X.done.exact(1));
  SLAM_CUDA_TRY(c, cudaMemcpy(X.peer_tab.p, X.peers.data(), sizeof(char*) * X.world, cudaMemcpyHostToDevice));
  SLAM_CUDA_TRY(c, cudaMemcpy(X.ranges.p, X.ranges_host.data(), sizeof(int) * 2 * X.world, cudaMemcpyHostToDevice));
  SLAM_CUDA_TRY(c, cudaMemset(X.err.p, 0, sizeof(int)));
  SLAM_CUDA_TRY(c, cudaMemset(X.done.p, 0, sizeof(unsigned)));
  X.epoch = 0;
  X.connected = true;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_assemble_exchange_async(slam_b200_ctx* c, int p0, int p1) try {
  NvtxRange nvtx_range("slam_b200/graph_assemble_exchange");
  if (!c || !c->sys) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  DeviceSystem& D = *c->sys;
  if (p0 < 0 || p1 > D.P || p0 > p1) return SLAM_B200_E_ARG;
  return graph_enqueue_assemble(c, p0, p1, false, true);
} SLAM_ABI_CATCH(c)

int slam_b200_xchg_set_timeout_ms(slam_b200_ctx* c, double ms) try {
  if (!c || !c->sys) return SLAM_B200_E_STATE;
  if (!(ms > 0.0) || ms > 3.6e6) return SLAM_B200_E_ARG;
  c->sys->xchg.timeout_ns = (unsigned long long)(ms * 1e6);
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_xchg_error(slam_b200_ctx* c) try {
  if (!c || !c->sys || !c->sys->xchg.connected) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  int e = 0;
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpy(&e, c->sys->xchg.err.p, sizeof(int), cudaMemcpyDeviceToHost));
  return e;
} SLAM_ABI_CATCH(c)

long slam_b200_graph_system_dev(slam_b200_ctx* c, int which, double** ptr) try {
  if (!c || !c->sys || !ptr) return SLAM_B200_E_STATE;
  DeviceSystem& D = *c->sys;
  *ptr = D.V.p;
  if (which == 0) return 6L * D.L;
  if (which == 1) return D.nV;
  if (which == 2) { *ptr = D.x.p; return D.n; }
  return SLAM_B200_E_ARG;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_solve_async(slam_b200_ctx* c) try {
  NvtxRange nvtx_range("slam_b200/graph_solve");
  if (!c || !c->sys || !c->sys->assembled || c->sys->assembly_only) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  return graph_enqueue_solve(c);
} SLAM_ABI_CATCH(c)

long slam_b200_graph_export_system(slam_b200_ctx* c, int* n_out, int32_t* Ap, int32_t* Ai, double* Ax, double* b) try {
  if (!c || !n_out) return SLAM_B200_E_ARG;
  int n = slam_b200_graph_prepare(c);
  if (n < 0) return n;
  DeviceSystem& D = *c->sys;
  *n_out = D.n;
  // scalar upper-triangular CSC in g2o order, built from the block list
  std::vector<std::map<int, std::pair<long, int>>> cols(D.n);  // col -> row -> (V offset, unused)
  auto dimOf = [&](int b) { return (D.blk_kind_local[b] & 1) ? 2 : 3; };
  for (int bb = 0; bb < D.nb; bb++) {
    int d = dimOf(bb), h = D.blk_hidx[bb];
    for (int i = 0; i < d; i++)
      for (int j = i; j < d; j++) cols[h + j][h + i] = {D.hoff_diag[bb] + i * d + j, 0};
  }
  for (size_t k = 0; k < D.off_a.size(); k++) {
    int a = D.off_a[k], bb = D.off_b[k];
    int da = dimOf(a), db = dimOf(bb), ha = D.blk_hidx[a], hb = D.blk_hidx[bb];
    for (int i = 0; i < da; i++)
      for (int j = 0; j < db; j++) cols[hb + j][ha + i] = {D.hoff_off[k] + i * db + j, 0};
  }
  long nnz = 0;
  for (auto& cm : cols) nnz += (long)cm.size();
  if (!Ai) return nnz;
  int rc = graph_enqueue_assemble(c, 0, D.P, false);
  if (rc) return rc;
  std::vector<double> V(D.nV);
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(V.data(), D.V.p, sizeof(double) * D.nV, cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  long q = 0;
  for (int j = 0; j < D.n; j++) {
    Ap[j] = (int)q;
    for (auto& kv : cols[j]) {
      Ai[q] = kv.first;
      Ax[q] = V[kv.second.first];
      q++;
    }
  }
  Ap[D.n] = (int)q;
  if (b)
    for (int bb = 0; bb < D.nb; bb++) {
      int kl = D.blk_kind_local[bb], h = D.blk_hidx[bb];
      if (kl & 1) {
        b[h] = V[2 * (size_t)(kl >> 1)];
        b[h + 1] = V[2 * (size_t)(kl >> 1) + 1];
      } else {
        for (int k = 0; k < 3; k++) b[h + k] = V[6 * (size_t)D.L + 3 * (size_t)(kl >> 1) + k];
      }
    }
  return nnz;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_stats(slam_b200_ctx* c, double out[16]) try {
  if (!c || !c->sys || !out) return SLAM_B200_E_STATE;
  DeviceSystem& D = *c->sys;
  for (int k = 0; k < 16; k++) out[k] = 0;
  out[0] = D.n; out[1] = D.nb; out[2] = D.sym.nf; out[3] = D.sym.nlevels;
  long nnzH = 0;
  for (int b = 0; b < D.nb; b++) { int d = (D.blk_kind_local[b] & 1) ? 2 : 3; nnzH += d * (d + 1) / 2; }
  for (size_t k = 0; k < D.off_a.size(); k++)
    nnzH += ((D.blk_kind_local[D.off_a[k]] & 1) ? 2 : 3) * ((D.blk_kind_local[D.off_b[k]] & 1) ? 2 : 3);
  out[4] = (double)nnzH; out[5] = (double)D.sym.nnzL; out[6] = D.sym.flops; out[7] = D.sym.max_front;
  out[8] = (double)(D.nL + D.nU); out[9] = D.sym.seconds; out[10] = D.upload_seconds;
  out[11] = (double)D.nV; out[12] = (double)D.nFbig; out[13] = (double)D.off_a.size();
  out[14] = D.t_structure; out[15] = D.t_lists;
  return 0;
} SLAM_ABI_CATCH(c)

long slam_b200_graph_export_symbolic(slam_b200_ctx* c, int what, int32_t* out, long cap);

// ---- device-side snapshot of the estimates (bench: every timed step starts from the same state) --
int slam_b200_graph_snapshot(slam_b200_ctx* c) try {
  if (!c || !c->sys || c->sys->R < 1) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  DeviceSystem& D = *c->sys;
  size_t n = (size_t)D.R * D.estStride;
  if (n) SLAM_CUDA_TRY(c, cudaMemcpyAsync(D.est0.p, D.est.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
  D.have_snapshot = true;
  return 0;
} SLAM_ABI_CATCH(c)
int slam_b200_graph_restore_async(slam_b200_ctx* c) try {
  if (!c || !c->sys || !c->sys->have_snapshot) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  DeviceSystem& D = *c->sys;
  size_t n = (size_t)D.R * D.estStride;
  if (n) SLAM_CUDA_TRY(c, cudaMemcpyAsync(D.est.p, D.est0.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
  SLAM_CUDA_TRY(c, cudaMemsetAsync(D.status.p, 0, sizeof(int) * 2 * (size_t)D.R, c->stream));
  D.iters_enqueued = 0;
  return 0;
} SLAM_ABI_CATCH(c)

// ---- per-phase timing with CUDA events (bench) ----------------------------------------------------
int slam_b200_profile_enable(slam_b200_ctx* c, int on) try {
  if (!c || !c->sys) return SLAM_B200_E_STATE;
  c->sys->profile = on != 0;
  return 0;
} SLAM_ABI_CATCH(c)
// out[0..4] = milliseconds summed over the profiled iterations: assemble, factor, forward, backward,
// update; out[5] = iterations profiled.  Synchronises the stream and clears the record.
int slam_b200_profile_read(slam_b200_ctx* c, double out[8]) try {
  if (!c || !c->sys || !out) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  DeviceSystem& D = *c->sys;
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 8; k++) out[k] = 0;
  size_t nit = D.prof_events.size() / 6;
  for (size_t it = 0; it < nit; it++)
    for (int ph = 0; ph < 5; ph++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, D.prof_events[6 * it + ph], D.prof_events[6 * it + ph + 1]);
      out[ph] += ms;
    }
  out[5] = (double)nit;
  for (cudaEvent_t e : D.prof_events) cudaEventDestroy(e);
  D.prof_events.clear();
  return 0;
} SLAM_ABI_CATCH(c)

// Debug: phase clocks (SM cycles) of block 0 of the most recent CTA-per-front factor launch; needs
// SLAM_B200_PHASE_CLOCKS=1 in the environment when the graph is prepared.  out[0..6] = clock64 at:
// start, zeroed, H scattered, children added, factorised, forward done, written; out[7..9] = s, fs, children.
int slam_b200_debug_phase_clocks(slam_b200_ctx* c, long long out[10]) try {
  if (!c || !c->sys || !out || !c->sys->dbg_clocks.p) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpy(out, c->sys->dbg_clocks.p, sizeof(long long) * 10, cudaMemcpyDeviceToHost));
  return 0;
} SLAM_ABI_CATCH(c)
// Debug: SM cycles summed over every warp of the warp-per-front factor kernel since the graph was
// prepared (same switch): out[0..5] = zero, scatter H, extend-add, LDL^T, fused forward, write; out[6] = fronts.
int slam_b200_debug_tiny_clocks(slam_b200_ctx* c, long long out[8]) try {
  if (!c || !c->sys || !out || !c->sys->dbg_clocks.p) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpy(out, c->sys->dbg_clocks.p + 16, sizeof(long long) * 8, cudaMemcpyDeviceToHost));
  return 0;
} SLAM_ABI_CATCH(c)

// Debug: cycles of thread 0 of the same CTA accumulated over the panels of its front (same switch):
// out[0..5] = triangle (warp 0), wait, row elimination, wait, trailing update, wait; out[6..7] = fused
// forward: warp-0 triangle solve, rest of the panel step.  Accumulates over launches until re-prepared.
int slam_b200_debug_panel_clocks(slam_b200_ctx* c, long long out[8]) try {
  if (!c || !c->sys || !out || !c->sys->dbg_clocks.p) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpy(out, c->sys->dbg_clocks.p + 24, sizeof(long long) * 8, cudaMemcpyDeviceToHost));
  SLAM_CUDA_TRY(c, cudaMemset(c->sys->dbg_clocks.p + 24, 0, sizeof(long long) * 8));
  return 0;
} SLAM_ABI_CATCH(c)

// Debug: front timeline (SLAM_B200_TIMELINE=1 at prepare time): six %globaltimer stamps (ns) per front of the last
// iteration -- factor kernel entry / after its wait / end, backward kernel entry / after its wait / end -- in the
// front numbering of slam_b200_graph_export_symbolic.  Returns the number of values (6 x fronts), copies min(cap, that).
long slam_b200_debug_timeline(slam_b200_ctx* c, long long* out, long cap) try {
  if (!c || !c->sys || !c->sys->timeline.p) return -1;
  if (ctx_set_device(c)) return -1;
  const long n = 12L * c->sys->sym.nf;
  if (!out) return n;
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
  if (cudaMemcpy(out, c->sys->timeline.p, sizeof(long long) * (size_t)std::min(n, cap), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return n;
} catch (...) { return -1; }

// ---- symbolic analysis without a device (host logic; testable on a CPU-only box) ----------------
struct SymHandle {
  Symbolic S;
  std::vector<int> hoff_diag, hoff_off;
};

void* slam_b200_symbolic_create(int nb, const int32_t* dim, int n_pairs, const int32_t* a, const int32_t* b,
                                int leaf_size) try {
  if (nb < 0 || n_pairs < 0 || (nb > 0 && !dim) || (n_pairs > 0 && (!a || !b))) return nullptr;
  std::unique_ptr<SymHandle> h(new SymHandle());
  int cur = 0;
  h->hoff_diag.resize(nb);
  for (int k = 0; k < nb; k++) { h->hoff_diag[k] = cur; cur += dim[k] * dim[k]; }
  h->hoff_off.resize(n_pairs);
  for (int k = 0; k < n_pairs; k++) { h->hoff_off[k] = cur; cur += dim[a[k]] * dim[b[k]]; }
  symbolic_analyze(nb, dim, n_pairs, a, b, h->hoff_diag.data(), h->hoff_off.data(), leaf_size, h->S);
  return h.release();
} catch (...) {
  return nullptr;  // no context to leave a message on
}
void slam_b200_symbolic_destroy(void* h) { delete static_cast<SymHandle*>(h); }

static long export_symbolic(const Symbolic& S, const std::vector<int>* blk_hidx, int what, int32_t* out, long cap) {
  std::vector<int> tmp;
  const std::vector<int>* v = nullptr;
  switch (what) {
    case 0: tmp = {S.nb, S.n, S.nf, S.nlevels, (int)S.upd_rows.size(), (int)S.asm_entries.size(), S.max_front}; v = &tmp; break;
    case 1: v = &S.pos; break;
    case 2: v = &S.boff; break;
    case 3: v = &S.level_ptr; break;
    case 4: v = &S.piv0; break;
    case 5: v = &S.npiv; break;
    case 6: v = &S.nupd; break;
    case 7: v = &S.parent; break;
    case 8: v = &S.rows_ptr; break;
    case 9: v = &S.upd_rows; break;
    case 10: v = &S.rel; break;
    case 11: v = &S.asm_ptr; break;
    case 12:
      for (auto& e : S.asm_entries) { tmp.push_back(e.hoff); tmp.push_back(e.r); tmp.push_back(e.c); tmp.push_back(e.meta); }
      v = &tmp;
      break;
    case 13: if (!blk_hidx) return SLAM_B200_E_ARG; v = blk_hidx; break;
    case 14: v = &S.dim; break;
    case 15: v = &S.child_ptr; break;
    case 16: v = &S.children; break;
    default: return SLAM_B200_E_ARG;
  }
  if (out) std::copy(v->begin(), v->begin() + std::min<long>((long)v->size(), cap), out);
  return (long)v->size();
}
long slam_b200_symbolic_export(void* h, int what, int32_t* out, long cap) {
  if (!h) return SLAM_B200_E_ARG;
  return export_symbolic(static_cast<SymHandle*>(h)->S, nullptr, what, out, cap);
}
// Tile plan of the batched factorisation (tileplan.h) for the analysed pattern, with the rhs of solver scalar k
// taken to sit at V offset nH + k (nH = number of H values): host logic, checked on the CPU by a numpy emulation of
// the kernel (tests/mf_emul.py).  what: 0 = {ok, nf, max_T, n_items, nF}, 1 = T, 2 = KT, 3 = fptr, 4 = item_ptr,
// 5 = item_nv, 6 = items as (src, dst) pairs.
long slam_b200_symbolic_tileplan(void* h, int what, int32_t* out, long cap) try {
  if (!h) return SLAM_B200_E_ARG;
  SymHandle* H = static_cast<SymHandle*>(h);
  const Symbolic& S = H->S;
  int nH = 0;
  for (int b = 0; b < S.nb; b++) nH += S.dim[b] * S.dim[b];
  {
    // off-diagonal blocks follow the diagonal ones (slam_b200_symbolic_create lays them out back to back)
    long last = nH;
    for (const AsmEntry& e : S.asm_entries) last = std::max(last, (long)e.hoff + (e.meta & 0xff) * ((e.meta >> 8) & 0xff));
    nH = (int)last;
  }
  std::vector<int> solver2v(S.n);
  for (int k = 0; k < S.n; k++) solver2v[k] = nH + k;
  TilePlan P;
  tile_plan_build(S, solver2v, P);
  std::vector<int> tmp;
  const std::vector<int>* v = &tmp;
  switch (what) {
    case 0: tmp = {P.ok ? 1 : 0, P.nf, P.max_T, (int)P.items.size(), (int)P.nF(), nH}; break;
    case 1: v = &P.T; break;
    case 2: v = &P.KT; break;
    case 3: for (long x : P.fptr) tmp.push_back((int)x); break;
    case 4: v = &P.item_ptr; break;
    case 5: v = &P.item_nv; break;
    case 6: for (const TileItem& it : P.items) { tmp.push_back(it.src); tmp.push_back(it.dst); } break;
    default: return SLAM_B200_E_ARG;
  }
  if (out) std::copy(v->begin(), v->begin() + std::min<long>((long)v->size(), cap), out);
  return (long)v->size();
} catch (...) { return SLAM_B200_E_NOMEM; }

double slam_b200_symbolic_stat(void* h, int what) {
  if (!h) return -1;
  const Symbolic& S = static_cast<SymHandle*>(h)->S;
  switch (what) {
    case 0: return (double)S.nnzL;
    case 1: return S.flops;
    case 2: return S.max_front;
    case 3: return S.seconds;
    case 4: return (double)(S.lptr[S.nf] + S.uptr[S.nf]);
    case 5: return S.t_nd;
    case 6: return S.t_md;
    default: return -1;
  }
}

long slam_b200_graph_export_symbolic(slam_b200_ctx* c, int what, int32_t* out, long cap) try {
  if (!c || !c->sys) return SLAM_B200_E_STATE;
  return export_symbolic(c->sys->sym, &c->sys->blk_hidx, what, out, cap);
} SLAM_ABI_CATCH(c)

// ------------------------------------------------------------------------------------------------
// batched replicas
// ------------------------------------------------------------------------------------------------
int slam_b200_batch_upload(slam_b200_ctx* c, int R, const double* pose_est3, const double* lm_est2,
                           const double* eo_z3, const double* el_z2) try {
  NvtxRange nvtx_range("slam_b200/batch_upload");
  if (!c || R < 1 || !pose_est3 || !lm_est2 || !eo_z3 || !el_z2) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  c->batch_ordering = true;
  int n = graph_build_structure(c);
  c->batch_ordering = false;
  if (n < 0) return n;
  DeviceSystem& D = *c->sys;
  int rc = graph_alloc_values(c, R);
  if (rc) return rc;
  const int P = D.P, L = D.L, Eo = D.Eo, El = D.El;
  const size_t ne = (size_t)D.estStride, nm = (size_t)D.measStride;
  // AoS (caller) -> SoA (device layout) through a pinned staging buffer, a few replicas at a time
  const int chunk = std::max(1, (int)std::min<size_t>((size_t)R, (64u << 20) / (8 * (ne + nm + 1))));
  SLAM_CUDA_TRY(c, c->pin_d.reserve((size_t)chunk * (ne + nm)));
  for (int r0 = 0; r0 < R; r0 += chunk) {
    int rn = std::min(chunk, R - r0);
    double* e = c->pin_d.p;
    double* m = e + (size_t)rn * ne;
    for (int rr = 0; rr < rn; rr++) {
      const int r = r0 + rr;
      double* er = e + (size_t)rr * ne;
      const double* pe = pose_est3 + (size_t)r * 3 * P;
      for (int p = 0; p < P; p++) { er[p] = pe[3 * p]; er[P + p] = pe[3 * p + 1]; er[2 * (size_t)P + p] = pe[3 * p + 2]; }
      const double* le = lm_est2 + (size_t)r * 2 * L;
      for (int l = 0; l < L; l++) { er[3 * (size_t)P + l] = le[2 * l]; er[3 * (size_t)P + L + l] = le[2 * l + 1]; }
      double* mr = m + (size_t)rr * nm;
      const double* lz = el_z2 + (size_t)r * 2 * El;
      for (int q = 0; q < El; q++) { int o = D.el_perm[q]; mr[q] = lz[2 * (size_t)o]; mr[El + (size_t)q] = lz[2 * (size_t)o + 1]; }
      const double* oz = eo_z3 + (size_t)r * 3 * Eo;
      for (int k = 0; k < Eo; k++) {
        mr[2 * (size_t)El + k] = oz[3 * (size_t)k];
        mr[2 * (size_t)El + Eo + k] = oz[3 * (size_t)k + 1];
        mr[2 * (size_t)El + 2 * (size_t)Eo + k] = oz[3 * (size_t)k + 2];
      }
      double* ml = mr + 2 * (size_t)El + 3 * (size_t)Eo;  // landmark order
      for (int q = 0; q < El; q++) { int e = D.lm_order[q]; ml[q] = mr[e]; ml[El + (size_t)q] = mr[El + (size_t)e]; }
    }
    if (ne) SLAM_CUDA_TRY(c, cudaMemcpyAsync(D.est.p + (size_t)r0 * ne, e, sizeof(double) * rn * ne, cudaMemcpyHostToDevice, c->stream));
    if (nm) SLAM_CUDA_TRY(c, cudaMemcpyAsync(D.meas.p + (size_t)r0 * nm, m, sizeof(double) * rn * nm, cudaMemcpyHostToDevice, c->stream));
    SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  }
  D.values_version = 0;  // replica 0 no longer mirrors the host graph
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_batch_iterate_async(slam_b200_ctx* c, int iters) { return slam_b200_graph_iterate_async(c, iters); }

int slam_b200_batch_download(slam_b200_ctx* c, double* pose_est3, double* lm_est2, double* chi2,
                             int chi2_cap_per_replica, int32_t* iterations_done) try {
  NvtxRange nvtx_range("slam_b200/batch_download");
  if (!c || !c->sys) return SLAM_B200_E_STATE;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  DeviceSystem& D = *c->sys;
  const int R = D.R, P = D.P, L = D.L;
  int rc = graph_enqueue_assemble(c, 0, D.P, true);
  if (rc) return rc;
  const size_t ne = (size_t)D.estStride;
  std::vector<int> st(2 * (size_t)R);
  std::vector<double> ch((size_t)R * D.chi2_cap);
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(st.data(), D.status.p, sizeof(int) * st.size(), cudaMemcpyDeviceToHost, c->stream));
  SLAM_CUDA_TRY(c, cudaMemcpyAsync(ch.data(), D.chi2.p, sizeof(double) * ch.size(), cudaMemcpyDeviceToHost, c->stream));
  const int chunk = std::max(1, (int)std::min<size_t>((size_t)R, (64u << 20) / (8 * (ne + 1))));
  SLAM_CUDA_TRY(c, c->pin_d.reserve((size_t)chunk * ne));
  for (int r0 = 0; r0 < R; r0 += chunk) {
    int rn = std::min(chunk, R - r0);
    SLAM_CUDA_TRY(c, cudaMemcpyAsync(c->pin_d.p, D.est.p + (size_t)r0 * ne, sizeof(double) * rn * ne, cudaMemcpyDeviceToHost, c->stream));
    SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int rr = 0; rr < rn; rr++) {
      const int r = r0 + rr;
      const double* er = c->pin_d.p + (size_t)rr * ne;
      if (pose_est3) {
        double* pe = pose_est3 + (size_t)r * 3 * P;
        for (int p = 0; p < P; p++) { pe[3 * p] = er[p]; pe[3 * p + 1] = er[P + p]; pe[3 * p + 2] = er[2 * (size_t)P + p]; }
      }
      if (lm_est2) {
        double* le = lm_est2 + (size_t)r * 2 * L;
        for (int l = 0; l < L; l++) { le[2 * l] = er[3 * (size_t)P + l]; le[2 * l + 1] = er[3 * (size_t)P + L + l]; }
      }
    }
  }
  SLAM_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int r = 0; r < R; r++) {
    int done = st[2 * r + 1], failed = st[2 * r];
    if (iterations_done) iterations_done[r] = failed ? 0 : done;
    if (chi2)
      for (int k = 0; k < done && k < chi2_cap_per_replica && k + 1 < D.chi2_cap; k++)
        chi2[(size_t)r * chi2_cap_per_replica + k] = ch[(size_t)r * D.chi2_cap + k + 1];
  }
  SLAM_CUDA_TRY(c, cudaMemsetAsync(D.status.p, 0, sizeof(int) * 2 * (size_t)R, c->stream));
  D.iters_enqueued = 0;
  return 0;
} SLAM_ABI_CATCH(c)

int slam_b200_graph_optimize_batch(slam_b200_ctx* c, int R, double* pose_est3, double* lm_est2,
                                   const double* eo_z3, const double* el_z2, int iters, double* chi2,
                                   int32_t* iterations_done) try {
  int rc = slam_b200_batch_upload(c, R, pose_est3, lm_est2, eo_z3, el_z2);
  if (rc) return rc;
  rc = slam_b200_batch_iterate_async(c, iters);
  if (rc) return rc;
  return slam_b200_batch_download(c, pose_est3, lm_est2, chi2, iters, iterations_done);
} SLAM_ABI_CATCH(c)

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// fp64 FMA peak of the device (the denominator for the solve's fp64 roofline; MEASURED_PEAKS.json
// only carries HBM bandwidth and bf16 tensor throughput)
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace

// Guard-band check (SLAM_B200_GUARD=1): re-reads the 0xFF bands in front of and behind every live device
// array of this process and counts the bytes that changed.  Returns that count (0 = no out-of-bounds
// write within 4 KiB of any array), -1 when the mode is off; *n_arrays = arrays checked.
extern "C" long slam_b200_debug_guard_check(slam_b200_ctx* c, long* n_arrays) try {
  if (!c) return SLAM_B200_E_ARG;
  if (!guard_mode()) return -1;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  SLAM_CUDA_TRY(c, cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_guard_mu);
  std::vector<unsigned char> h(2 * GUARD_BYTES + 256);
  long bad = 0, n = 0;
  for (const auto& kv : guard_tab()) {
    const char* raw = static_cast<const char*>(kv.first);
    const size_t bytes = kv.second, padded = (bytes + 255) & ~(size_t)255;
    SLAM_CUDA_TRY(c, cudaMemcpy(h.data(), raw, GUARD_BYTES, cudaMemcpyDeviceToHost));
    // behind the payload: the alignment padding belongs to the band too
    const size_t tail = GUARD_BYTES + (padded - bytes);
    SLAM_CUDA_TRY(c, cudaMemcpy(h.data() + GUARD_BYTES, raw + GUARD_BYTES + bytes, tail, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < GUARD_BYTES + tail; k++) bad += h[k] != 0xFF;
    n++;
  }
  if (n_arrays) *n_arrays = n;
  return bad;
} SLAM_ABI_CATCH(c)

extern "C" int slam_b200_fp64_peak(slam_b200_ctx* c, double* tflops) try {
  if (!c || !tflops) return SLAM_B200_E_ARG;
  if (ctx_set_device(c)) return SLAM_B200_E_CUDA;
  const int blocks = c->num_sms * 8, threads = 256, iters = 1 << 15;
  double* buf = nullptr;
  SLAM_CUDA_TRY(c, cudaMalloc(&buf, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0, c->stream);
    fp64_peak_kernel<<<blocks, threads, 0, c->stream>>>(buf, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 8 * (double)iters * blocks * threads;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  c->launches += 4;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  SLAM_CUDA_TRY(c, cudaGetLastError());
  *tflops = best;
  return 0;
} SLAM_ABI_CATCH(c)

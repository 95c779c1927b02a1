// oracle/slam_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the GraphSLAM back-end hot path of
// cfsd/opendlv-logic-cfsd18-sensation-slam.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product
// (the CUDA library under opendlv-logic-cfsd18-sensation-slam_b200/) never does.
//
// PARITY UNPINNED: the reference ships no golden vectors for this path (its only test is
// `5+6==11`, test/tests-logic-cfsd18-sensation-slam.cpp:26-30) and g2o -- which holds the
// Gauss-Newton arithmetic -- is neither vendored nor installed (Dockerfile.amd64:33 clones an
// unpinned HEAD).  The association / conversion half follows src/slam.cpp line by line; the
// graph half restates the published g2o algorithm (2018-era API, see the per-function notes)
// and anchors on the reference's call sites (slam.cpp:53-65, 433-484, 525-550, 713-732).
// It is pinned only by (i) the survey-derived known answers in tests/golden/known_answers.json,
// (ii) finite-difference / scipy cross-checks in tests/, and (iii) when built with
// -DORACLE_USE_EIGEN, by running the *reference's own vendored* Eigen 3.3.4
// SimplicialLDLT<Upper>+AMD (thirdparty/Eigen/src/SparseCholesky/SimplicialCholesky.h:156-238),
// the solver g2o's LinearSolverEigen wraps.
// Pinned by reference code proper (tests/test_pinned_by_reference.py): (a) the whole association /
// conversion / graph-building half -- oracle/build_ref_slam.sh compiles the reference's REAL
// src/slam.cpp with g2o replaced by oracle/g2o_facade (over the Gauss-Newton below) and the C1 replay
// through it (tests/golden/c1_replay_reference.npz) is reproduced identically by orc_slam_perform;
// (b) the WGS84 <-> Cartesian conversion (orc_wgs84_*), bit for bit against vectors produced by
// compiling src/WGS84toCartesian.hpp itself.  What stays unpinned is the g2o arithmetic.
//
// Build (see oracle/Makefile): g++ -std=c++14 -O2 -ffp-contract=off, no -march=native, i.e. the
// reference's own flags (CMakeLists.txt:35-38) so nothing is contracted into FMA.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <utility>
#include <vector>

#ifdef ORACLE_USE_EIGEN
#include <Eigen/Sparse>
#include <Eigen/SparseCholesky>
#endif

namespace {

// slam.hpp:134-136.  PI is a float literal widened to double (= 3.1415927410125732).
const double DEG2RAD = 0.017453292522222;
const double RAD2DEG = 57.295779513082325;
const double PI_REF = 3.14159265f;

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// slam.cpp:513-523
void transformConeToCoG(double angle, double distance, double out[2]) {
  const double lidarDistToCoG = 1.5;
  double sign = angle / std::fabs(angle);
  angle = PI_REF - std::fabs(angle * DEG2RAD);
  double distanceNew = std::sqrt(lidarDistToCoG * lidarDistToCoG + distance * distance -
                                 2 * lidarDistToCoG * distance * std::cos(angle));
  double angleNew = std::asin((std::sin(angle) * distance) / distanceNew) * RAD2DEG;
  out[0] = angleNew * sign;
  out[1] = distanceNew;
}

// slam.cpp:637-654
void spherical2Cartesian(double azimuth, double zenimuth, double distance, double out[3]) {
  double t[2];
  transformConeToCoG(azimuth, distance, t);
  azimuth = t[0];
  distance = t[1];
  out[0] = distance * std::cos(zenimuth * DEG2RAD) * std::cos(azimuth * DEG2RAD);
  out[1] = distance * std::cos(zenimuth * DEG2RAD) * std::sin(azimuth * DEG2RAD);
  out[2] = distance * std::sin(zenimuth * DEG2RAD);
}

// slam.cpp:499-510.  obs = one column (az, zen, range, type) of the 4xN frame matrix.
void coneToGlobal(const double pose[3], const double obs[4], double out[3]) {
  double c[3];
  spherical2Cartesian(obs[0], obs[1], obs[2], c);
  double newX = c[0] * std::cos(pose[2]) - c[1] * std::sin(pose[2]);
  double newY = c[0] * std::sin(pose[2]) + c[1] * std::cos(pose[2]);
  out[0] = newX + pose[0];
  out[1] = newY + pose[1];
  out[2] = obs[3];
}

// slam.cpp:708-711
inline double distanceBetweenCones(double x1, double y1, double x2, double y2) {
  return std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// conversions (slam.cpp:499-523, 637-654)
// ---------------------------------------------------------------------------------------------
void orc_transform_cone_to_cog(double angle, double distance, double* out2) {
  transformConeToCoG(angle, distance, out2);
}
void orc_spherical2cartesian(double az, double zen, double d, double* out3) {
  spherical2Cartesian(az, zen, d, out3);
}
void orc_cone_to_global(const double* pose3, const double* obs4, double* out3) {
  coneToGlobal(pose3, obs4, out3);
}
double orc_pi_ref() { return PI_REF; }

// ---------------------------------------------------------------------------------------------
// Mapping-phase association of one frame: Slam::addConesToMap, slam.cpp:552-635, minus the graph
// side effects (returned as per-observation records instead) and minus the optimise call itself
// (the caller runs optimise once per observation i >= *loop_closing_obs, slam.cpp:625-633).
//
// status[i]: 0 matched map cone idx[i] (edge added, 591); 1 new cone idx[i] created (610-619);
//            2 unmatched and not added (range >= coneMappingThreshold, 608); 3 skipped because
//            m_loopClosing was already set (575, 608).
// z[2*i..] : vehicle-frame xy measurement the edge carries (addConeMeasurement 539-545).
// g[3*i..] : coneToGlobal(pose, col i)  (572).
// first_cone_created: 1 if the map was empty and cone 0 was created from column 0 (554-567);
//            that cone's edge precedes all per-observation edges.
// Map arrays have capacity map_cap; *M is updated.  Returns 0, or -1 when capacity is exceeded.
// ---------------------------------------------------------------------------------------------
int orc_assoc_map_frame(const double* cones4xN, int N, const double* pose3, double sameConeThr,
                        double mapThr, double* map_x, double* map_y, int* map_type, int* M,
                        int map_cap, unsigned* current_cone_index, int* loop_closing,
                        int* idx, int* status, double* z2, double* g3, int* first_cone_created,
                        int* loop_closing_obs) {
  int m = *M;
  *first_cone_created = 0;
  *loop_closing_obs = -1;
  if (m == 0 && N > 0) {  // 554-567
    double g[3];
    coneToGlobal(pose3, cones4xN, g);
    if (m >= map_cap) return -1;
    map_x[0] = g[0];
    map_y[0] = g[1];
    map_type[0] = (int)g[2];
    m = 1;
    *first_cone_created = 1;
  }
  double minDistance = 100;  // 569
  for (int i = 0; i < N; i++) {
    const double* col = cones4xN + 4 * (size_t)i;
    double distanceToCar = col[2];
    double g[3];
    coneToGlobal(pose3, col, g);
    g3[3 * i] = g[0];
    g3[3 * i + 1] = g[1];
    g3[3 * i + 2] = g[2];
    double s[3];
    spherical2Cartesian(col[0], col[1], col[2], s);
    z2[2 * i] = s[0];
    z2[2 * i + 1] = s[1];
    idx[i] = -1;
    status[i] = *loop_closing ? 3 : 2;
    int j = 0;
    bool coneFound = false;
    while (!coneFound && j < m && !*loop_closing) {      // 575
      if (std::fabs(map_type[j] - col[3]) < 0.0001) {    // 576
        double distance = distanceBetweenCones(map_x[j], map_y[j], g[0], g[1]);  // 579
        if (distance < sameConeThr) {                    // 584
          coneFound = true;
          idx[i] = j;
          status[i] = 0;
          // loopClosing(m_map[j], distanceToCar), 697-706
          double lc = distanceBetweenCones(map_x[0], map_y[0], map_x[j], map_y[j]);
          if (lc < 1 && *current_cone_index > 20 && distanceToCar < mapThr && !*loop_closing) {
            *loop_closing = 1;                           // 595
            *loop_closing_obs = i;
          }
          if (distanceToCar < minDistance) {             // 598-601
            *current_cone_index = (unsigned)j;
            minDistance = distanceToCar;
          }
        }
      }
      j++;
    }
    if (distanceToCar < mapThr && !coneFound && !*loop_closing) {  // 608
      if (m >= map_cap) return -1;
      map_x[m] = g[0];
      map_y[m] = g[1];
      map_type[m] = (int)g[2];
      idx[i] = m;
      status[i] = 1;
      m++;
    }
  }
  *M = m;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Localisation-phase association of one frame: Slam::localizer, slam.cpp:340-387.
// Gate (360): dist < thr && (mapType - (int)obsType) < 0.0001   -- asymmetric, no fabs.
// idx[i] = first map index passing the gate, else -1.  *current_cone_index is updated as at 387;
// *send_cone_data (385) is only meaningful when *n_reobserved > 0 (the reference reads an
// uninitialised local otherwise, slam.cpp:348).
// ---------------------------------------------------------------------------------------------
int orc_assoc_localize_frame(const double* cones4xN, int N, const double* pose3, double thr,
                             const double* map_x, const double* map_y, const int* map_type, int M,
                             unsigned* current_cone_index, int* idx, double* g3, int* n_reobserved,
                             int* send_cone_data) {
  double minDistance = 100;
  unsigned cci = 0;
  int reobs = 0;
  for (int i = 0; i < N; i++) {
    const double* col = cones4xN + 4 * (size_t)i;
    double g[3];
    coneToGlobal(pose3, col, g);
    g3[3 * i] = g[0];
    g3[3 * i + 1] = g[1];
    g3[3 * i + 2] = g[2];
    double distanceToCar = col[2];
    int obsType = static_cast<int>(g[2]);  // 358
    idx[i] = -1;
    for (int j = 0; j < M; j++) {
      if (distanceBetweenCones(map_x[j], map_y[j], g[0], g[1]) < thr &&
          (map_type[j] - obsType) < 0.0001) {  // 360
        reobs++;
        idx[i] = j;
        if (distanceToCar < minDistance) {  // 375-378
          cci = (unsigned)j;
          minDistance = distanceToCar;
        }
        break;
      }
    }
  }
  *n_reobserved = reobs;
  *send_cone_data = (reobs > 0) ? (cci != *current_cone_index) : 0;
  if (reobs > 0) *current_cone_index = cci;  // 387
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Match-only association against a frozen map (phase 1 of addConesToMap / the localizer loop),
// used for the large-field configuration.  mode 0: mapping gate (576, 584); mode 1: localizer
// gate (360).  min_margin (optional) returns min over all tested same-type pairs of
// |dist - thr| so tests can prove no decision sits on a rounding knife-edge.
// ---------------------------------------------------------------------------------------------
int orc_assoc_match_only(const double* cones4xN, int N, const double* pose3, double thr, int mode,
                         const double* map_x, const double* map_y, const int* map_type, int M,
                         int* idx, double* g3, double* min_margin) {
  double mm = 1e300;
  for (int i = 0; i < N; i++) {
    const double* col = cones4xN + 4 * (size_t)i;
    double g[3];
    coneToGlobal(pose3, col, g);
    if (g3) {
      g3[3 * i] = g[0];
      g3[3 * i + 1] = g[1];
      g3[3 * i + 2] = g[2];
    }
    int obsType = (mode == 1) ? static_cast<int>(g[2]) : 0;
    int found = -1;
    for (int j = 0; j < M; j++) {
      bool typeOk = (mode == 0) ? (std::fabs(map_type[j] - col[3]) < 0.0001)
                                : ((map_type[j] - obsType) < 0.0001);
      if (!typeOk) continue;
      double d = distanceBetweenCones(map_x[j], map_y[j], g[0], g[1]);
      double mg = std::fabs(d - thr);
      if (mg < mm) mm = mg;
      if (d < thr) {
        found = j;
        break;
      }
    }
    idx[i] = found;
  }
  if (min_margin) *min_margin = mm;
  return 0;
}

}  // extern "C"

// =============================================================================================
// GNSS priors: WGS84 -> local Cartesian as Slam::nextPose / nextSplitPose apply it (slam.cpp:153-209)
// and back as Slam::sendPose does (681-695), restated from the reference's vendored
// src/WGS84toCartesian.hpp (toCartesian 40-118, fromCartesian 125-155).  This is the one row of the
// path whose arithmetic lives in /root/reference itself, so it IS pinned by reference code:
// tests/golden/wgs84_vectors.json is generated by compiling that header (make_wgs84_golden.py).
// =============================================================================================
namespace wgs84_restated {
const double PI = 3.14159265358979323846, D2R = PI / 180.0, A = 6378137.0;
const double F = 1.0 / 298.257223563, ES = 2.0 * F - F * F;

// meridional arc (unit ellipsoid), series coefficients of WGS84toCartesian.hpp:52-72
static double arc(double lat) {
  const double q = ES;
  const double R0 = 1.0 - q * (0.25 + q * (0.046875 + q * (0.01953125 + q * 0.01068115234375)));
  const double R1 = q * (0.75 - q * (0.046875 + q * (0.01953125 + q * 0.01068115234375)));
  const double R2T = q * q;
  const double R2 = R2T * (0.46875 - q * (0.01302083333333333333 + q * 0.00712076822916666666));
  const double R3T = R2T * q;
  const double R3 = R3T * (0.36458333333333333333 - q * 0.00569661458333333333);
  const double R4 = R3T * q * 0.3076171875;
  const double s = std::sin(lat), cs = std::cos(lat) * s, s2 = s * s;
  return R0 * lat - cs * (R1 + s2 * (R2 + s2 * (R3 + s2 * R4)));
}

static void to_cartesian(const double* ref, const double* pos, double* out) {
  double lat = pos[0] * D2R, lon = pos[1] * D2R;
  const double ml0 = arc(ref[0] * D2R);
  const double over = std::abs(lat) - PI / 2.0;
  if (over > 1.0e-12 || std::abs(lon) > 10.0) { out[0] = out[1] = 0.0; return; }   // :103-105
  if (std::abs(over) < 1.0e-12) lat = lat < 0.0 ? -1.0 * (PI / 2.0) : PI / 2.0;
  lon -= ref[1] * D2R;
  double x = lon, y = -1.0 * ml0;                                                    // :86
  if (!(std::abs(lat) < 1.0e-10)) {
    const double sl = std::sin(lat);
    const double ms = std::abs(sl) > 1.0e-10 ? (std::cos(lat) / std::sqrt(1.0 - ES * sl * sl)) / sl : 0.0;
    lon *= sl;
    x = ms * std::sin(lon);
    y = (arc(lat) - ml0) + ms * (1.0 - std::cos(lon));
  }
  out[0] = A * x;
  out[1] = A * y;
}

static void from_cartesian(const double* ref, const double* xy, double* out) {
  double g[2] = {ref[0], ref[1]}, c[2];
  to_cartesian(ref, g, c);
  const int sLat = xy[1] < 0 ? -1 : 1, sLon = xy[0] < 0 ? -1 : 1;
  double prev = std::numeric_limits<double>::max(), d = std::abs(xy[1] - c[1]);
  while (d < prev && d > 1.0e-2) {          // latitude against the y residual (:137-143)
    g[0] = g[0] + sLat * 1e-5;
    to_cartesian(ref, g, c);
    prev = d;
    d = std::abs(xy[1] - c[1]);
  }
  prev = std::numeric_limits<double>::max();
  d = std::abs(xy[0] - c[0]);
  while (d < prev && d > 1.0e-2) {          // longitude against the x residual (:145-152)
    g[1] = g[1] + sLon * 1e-5;
    to_cartesian(ref, g, c);
    prev = d;
    d = std::abs(xy[0] - c[0]);
  }
  out[0] = g[0];
  out[1] = g[1];
}
}  // namespace wgs84_restated

extern "C" {
void orc_wgs84_to_cartesian(const double* ref2, const double* pos2, double* out2) { wgs84_restated::to_cartesian(ref2, pos2, out2); }
void orc_wgs84_from_cartesian(const double* ref2, const double* xy2, double* out2) { wgs84_restated::from_cartesian(ref2, xy2, out2); }
// Slam::nextSplitPose (slam.cpp:176-181): GeodeticHeadingReading -> heading about PI (PI = 3.14159265f)
double orc_heading_from_north(float northHeading) {
  double heading = northHeading;
  heading = heading - PI_REF;
  heading = (heading > PI_REF) ? (heading - 2 * PI_REF) : (heading);
  heading = (heading < -PI_REF) ? (heading + 2 * PI_REF) : (heading);
  return heading;
}
}

// =============================================================================================
// mini-g2o: the Gauss-Newton back end the reference reaches through g2o (slam.hpp:26-35).
// Restated from the published g2o sources of the 2018 era (github.com/RainerKuemmerle/g2o,
// tag 20200410_git is API-compatible): types/slam2d/{se2.h,vertex_se2.h,vertex_point_xy.h,
// edge_se2.cpp,edge_se2_pointxy.cpp}, stuff/misc.h, core/{sparse_optimizer.cpp,
// optimization_algorithm_gauss_newton.cpp,block_solver.hpp,base_binary_edge.hpp},
// solvers/eigen/linear_solver_eigen.h.
// =============================================================================================
namespace {

// g2o stuff/misc.h normalize_theta (2018 form)
inline double normalize_theta(double theta) {
  if (theta >= -M_PI && theta < M_PI) return theta;
  double multiplier = std::floor(theta / (2 * M_PI));
  theta = theta - multiplier * 2 * M_PI;
  if (theta >= M_PI) theta -= 2 * M_PI;
  if (theta < -M_PI) theta += 2 * M_PI;
  return theta;
}

// g2o types/slam2d/se2.h
struct SE2 {
  double x, y, th;
  SE2() : x(0), y(0), th(0) {}
  SE2(double x_, double y_, double th_) : x(x_), y(y_), th(th_) {}
  // Rotation2D * v evaluates sin/cos of the stored angle (Eigen Geometry/Rotation2D.h:109-110,188)
  void rot(double vx, double vy, double& ox, double& oy) const {
    double s = std::sin(th), c = std::cos(th);
    ox = c * vx - s * vy;
    oy = s * vx + c * vy;
  }
  SE2 operator*(const SE2& b) const {  // operator*=: _t += _R*b._t; angle += b.angle; normalize
    SE2 r(*this);
    double rx, ry;
    rot(b.x, b.y, rx, ry);
    r.x = x + rx;
    r.y = y + ry;
    r.th = normalize_theta(th + b.th);
    return r;
  }
  void mulPoint(double vx, double vy, double& ox, double& oy) const {  // _t + _R*v
    double rx, ry;
    rot(vx, vy, rx, ry);
    ox = x + rx;
    oy = y + ry;
  }
  SE2 inverse() const {
    SE2 r;
    r.th = normalize_theta(-th);
    r.rot(-1 * x, -1 * y, r.x, r.y);
    return r;
  }
};

struct Vertex {
  int id;
  int dim;  // 3: VertexSE2, 2: VertexPointXY
  double est[3];
  bool fixed;
  int hidx;  // scalar offset in the Hessian, -1 if fixed / inactive
  int bidx;  // block index
};

struct Edge {
  int kind;  // 0: EdgeSE2 (pose-pose), 1: EdgeSE2PointXY (pose-landmark)
  int vi, vj;
  double z[3];
  double info[9];  // row-major DxD
  SE2 zinv;        // EdgeSE2::_inverseMeasurement
};

struct Block {
  int r, c;       // block row / column (r <= c, upper)
  int dr, dc;
  double v[9];    // row-major dr x dc
};

struct Graph {
  std::vector<Vertex> verts;
  std::map<int, int> id2v;
  std::vector<Edge> edges;
  // timings of the last optimize()
  double t_init = 0, t_struct = 0, t_linearize = 0, t_analyze = 0, t_factor = 0, t_solve = 0,
         t_update = 0, t_total = 0;
  long nnzH = 0, nnzL = 0;
  double flopsL = 0;
  int nfree = 0;
};

// EdgeSE2::computeError / EdgeSE2PointXY::computeError
void computeError(const Graph& G, const Edge& e, double err[3]) {
  const Vertex& a = G.verts[e.vi];
  const Vertex& b = G.verts[e.vj];
  if (e.kind == 0) {
    SE2 vi(a.est[0], a.est[1], a.est[2]), vj(b.est[0], b.est[1], b.est[2]);
    SE2 delta = e.zinv * (vi.inverse() * vj);
    err[0] = delta.x;
    err[1] = delta.y;
    err[2] = delta.th;
  } else {
    SE2 vi(a.est[0], a.est[1], a.est[2]);
    double px, py;
    vi.inverse().mulPoint(b.est[0], b.est[1], px, py);
    err[0] = px - e.z[0];
    err[1] = py - e.z[1];
    err[2] = 0;
  }
}

inline int edgeDim(const Edge& e) { return e.kind == 0 ? 3 : 2; }

double edgeChi2(const Graph& G, const Edge& e) {
  double err[3];
  computeError(G, e, err);
  int D = edgeDim(e);
  double c = 0;
  for (int r = 0; r < D; r++) {
    double t = 0;
    for (int k = 0; k < D; k++) t += e.info[r * D + k] * err[k];
    c += err[r] * t;
  }
  return c;
}

// analytic Jacobians: edge_se2.cpp / edge_se2_pointxy.cpp linearizeOplus.
// Ji is D x 3, Jj is D x dj (row-major, leading dimension 3).
void linearize(const Graph& G, const Edge& e, double Ji[9], double Jj[9]) {
  const Vertex& a = G.verts[e.vi];
  const Vertex& b = G.verts[e.vj];
  for (int k = 0; k < 9; k++) Ji[k] = Jj[k] = 0;
  if (e.kind == 0) {
    double thetai = a.est[2];
    double dtx = b.est[0] - a.est[0], dty = b.est[1] - a.est[1];
    double si = std::sin(thetai), ci = std::cos(thetai);
    double A[9] = {-ci, -si, -si * dtx + ci * dty, si, -ci, -ci * dtx - si * dty, 0, 0, -1};
    double B[9] = {ci, si, 0, -si, ci, 0, 0, 0, 1};
    double s = std::sin(e.zinv.th), c = std::cos(e.zinv.th);
    double Z[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
    for (int r = 0; r < 3; r++)
      for (int cc = 0; cc < 3; cc++) {
        double sa = 0, sb = 0;
        for (int k = 0; k < 3; k++) {
          sa += Z[r * 3 + k] * A[k * 3 + cc];
          sb += Z[r * 3 + k] * B[k * 3 + cc];
        }
        Ji[r * 3 + cc] = sa;
        Jj[r * 3 + cc] = sb;
      }
  } else {
    double x1 = a.est[0], y1 = a.est[1], th1 = a.est[2];
    double x2 = b.est[0], y2 = b.est[1];
    double aux_1 = std::cos(th1), aux_2 = -aux_1, aux_3 = std::sin(th1);
    Ji[0] = aux_2;
    Ji[1] = -aux_3;
    Ji[2] = aux_1 * y2 - aux_1 * y1 - aux_3 * x2 + aux_3 * x1;
    Ji[3] = aux_3;
    Ji[4] = aux_2;
    Ji[5] = -aux_3 * y2 + aux_3 * y1 - aux_1 * x2 + aux_1 * x1;
    Jj[0] = aux_1;
    Jj[1] = aux_3;
    Jj[3] = -aux_3;
    Jj[4] = aux_1;
  }
}

// VertexSE2::oplusImpl (additive, normalised angle) / VertexPointXY::oplusImpl
void oplus(Vertex& v, const double* u) {
  if (v.dim == 3) {
    v.est[0] += u[0];
    v.est[1] += u[1];
    v.est[2] = normalize_theta(v.est[2] + u[2]);
  } else {
    v.est[0] += u[0];
    v.est[1] += u[1];
  }
}

// ---------------------------------------------------------------------------------------------
// "port" linear solver: up-looking simplicial LDL^T, restating the algorithm of
// thirdparty/Eigen/src/SparseCholesky/SimplicialCholesky_impl.h:51-96 (etree + column counts)
// and :101-195 (numeric), on the upper triangle of P*A*P^T.  The fill-reducing permutation is a
// plain exact minimum-degree on the scalar pattern (Eigen uses AMD, OrderingMethods/Amd.h:94;
// any permutation gives the same x up to rounding).  Used when the library is built without
// the vendored Eigen.
// ---------------------------------------------------------------------------------------------
struct CscUpper {
  int n = 0;
  std::vector<int> Ap, Ai;  // column pointers / row indices (rows <= col), sorted
  std::vector<double> Ax;
};

std::vector<int> minimumDegreeOrder(const CscUpper& A) {
  int n = A.n;
  std::vector<std::set<int>> adj(n);
  for (int j = 0; j < n; j++)
    for (int p = A.Ap[j]; p < A.Ap[j + 1]; p++) {
      int i = A.Ai[p];
      if (i != j) {
        adj[i].insert(j);
        adj[j].insert(i);
      }
    }
  std::set<std::pair<int, int>> pq;
  for (int v = 0; v < n; v++) pq.insert({(int)adj[v].size(), v});
  std::vector<int> perm;
  perm.reserve(n);
  std::vector<char> done(n, 0);
  while (!pq.empty()) {
    int v = pq.begin()->second;
    pq.erase(pq.begin());
    done[v] = 1;
    perm.push_back(v);
    std::vector<int> nb(adj[v].begin(), adj[v].end());
    for (int u : nb) {
      pq.erase({(int)adj[u].size(), u});
      adj[u].erase(v);
    }
    for (size_t a = 0; a < nb.size(); a++)
      for (size_t b = a + 1; b < nb.size(); b++) {
        adj[nb[a]].insert(nb[b]);
        adj[nb[b]].insert(nb[a]);
      }
    for (int u : nb) pq.insert({(int)adj[u].size(), u});
    adj[v].clear();
  }
  return perm;  // perm[k] = original index eliminated k-th
}

struct SimplicialLDLT {
  int n = 0;
  std::vector<int> P, Pinv, parent, Lp, Li, nzCount;
  std::vector<double> Lx, D;
  // permuted upper pattern
  std::vector<int> Cp, Ci, Cmap;  // Cmap: position in A.Ax for each permuted entry
  std::vector<double> Cx;
  bool ok = false;
  double flops = 0;

  void analyze(const CscUpper& A) {
    n = A.n;
    P = minimumDegreeOrder(A);
    Pinv.assign(n, 0);
    for (int k = 0; k < n; k++) Pinv[P[k]] = k;
    // C = upper(P A P^T): entry (i,j) of A -> (min,max) of (Pinv[i],Pinv[j])
    std::vector<int> cnt(n + 1, 0);
    for (int j = 0; j < n; j++)
      for (int p = A.Ap[j]; p < A.Ap[j + 1]; p++) {
        int a = Pinv[A.Ai[p]], b = Pinv[j];
        cnt[std::max(a, b) + 1]++;
      }
    Cp.assign(n + 1, 0);
    for (int j = 0; j < n; j++) Cp[j + 1] = Cp[j] + cnt[j + 1];
    Ci.assign(Cp[n], 0);
    Cmap.assign(Cp[n], 0);
    std::vector<int> w(Cp.begin(), Cp.end() - 1);
    for (int j = 0; j < n; j++)
      for (int p = A.Ap[j]; p < A.Ap[j + 1]; p++) {
        int a = Pinv[A.Ai[p]], b = Pinv[j];
        int q = w[std::max(a, b)]++;
        Ci[q] = std::min(a, b);
        Cmap[q] = p;
      }
    Cx.assign(Cp[n], 0);
    // etree + column counts (_impl.h:51-96)
    parent.assign(n, -1);
    nzCount.assign(n, 0);
    std::vector<int> tags(n, 0);
    for (int k = 0; k < n; k++) {
      parent[k] = -1;
      tags[k] = k;
      nzCount[k] = 0;
      for (int p = Cp[k]; p < Cp[k + 1]; p++) {
        int i = Ci[p];
        if (i < k) {
          for (; tags[i] != k; i = parent[i]) {
            if (parent[i] == -1) parent[i] = k;
            nzCount[i]++;
            tags[i] = k;
          }
        }
      }
    }
    Lp.assign(n + 1, 0);
    flops = 0;
    for (int k = 0; k < n; k++) {
      Lp[k + 1] = Lp[k] + nzCount[k];
      double c = nzCount[k];
      flops += c * c + 3 * c;
    }
    Li.assign(Lp[n], 0);
    Lx.assign(Lp[n], 0);
    D.assign(n, 0);
  }

  // _impl.h:101-195
  bool factorize(const CscUpper& A) {
    for (size_t q = 0; q < Cmap.size(); q++) Cx[q] = A.Ax[Cmap[q]];
    std::vector<double> y(n, 0);
    std::vector<int> pattern(n, 0), tags(n, 0);
    ok = true;
    for (int k = 0; k < n; k++) {
      y[k] = 0;
      int top = n;
      tags[k] = k;
      nzCount[k] = 0;
      for (int p = Cp[k]; p < Cp[k + 1]; p++) {
        int i = Ci[p];
        if (i <= k) {
          y[i] += Cx[p];
          int len;
          for (len = 0; tags[i] != k; i = parent[i]) {
            pattern[len++] = i;
            tags[i] = k;
          }
          while (len > 0) pattern[--top] = pattern[--len];
        }
      }
      double d = y[k];
      y[k] = 0;
      for (; top < n; ++top) {
        int i = pattern[top];
        double yi = y[i];
        y[i] = 0;
        double l_ki = yi / D[i];
        int p2 = Lp[i] + nzCount[i];
        int p;
        for (p = Lp[i]; p < p2; ++p) y[Li[p]] -= Lx[p] * yi;
        d -= l_ki * yi;
        Li[p] = k;
        Lx[p] = l_ki;
        ++nzCount[i];
      }
      D[k] = d;
      if (d == 0) {  // _impl.h:175-179
        ok = false;
        return false;
      }
    }
    return true;
  }

  // SimplicialCholesky.h:156-180: x = Pinv L^-T D^-1 L^-1 P b
  void solve(const double* b, double* x) const {
    std::vector<double> y(n);
    for (int k = 0; k < n; k++) y[k] = b[P[k]];
    for (int j = 0; j < n; j++)
      for (int p = Lp[j]; p < Lp[j + 1]; p++) y[Li[p]] -= Lx[p] * y[j];
    for (int k = 0; k < n; k++) y[k] /= D[k];
    for (int j = n - 1; j >= 0; j--)
      for (int p = Lp[j]; p < Lp[j + 1]; p++) y[j] -= Lx[p] * y[Li[p]];
    for (int k = 0; k < n; k++) x[P[k]] = y[k];
  }
};

// ---------------------------------------------------------------------------------------------
// The optimiser proper
// ---------------------------------------------------------------------------------------------
struct System {
  std::vector<int> activeEdges;           // by insertion id
  std::vector<int> ivMap;                 // block index -> vertex index (non-fixed, id order)
  std::map<std::pair<int, int>, int> blockOf;  // (block row, block col), r<=c -> index in blocks
  std::vector<Block> blocks;
  std::vector<double> b, x;
  CscUpper A;
  std::vector<int> blockCsc;  // for each block, for each scalar (row-major) -> index in A.Ax or -1
  // per active edge: block ids of H_ii, H_jj, H_ij (or -1), like the Hessian-block pointers g2o
  // hands every edge in BlockSolver::buildStructure
  std::vector<int> edgeBlk;
  int n = 0;
};

// SparseOptimizer::initializeOptimization + buildIndexMapping
bool initializeOptimization(Graph& G, System& S) {
  S.activeEdges.clear();
  std::vector<char> activeV(G.verts.size(), 0);
  for (size_t k = 0; k < G.edges.size(); k++) {
    const Edge& e = G.edges[k];
    bool allFixed = G.verts[e.vi].fixed && G.verts[e.vj].fixed;
    if (!allFixed) {
      S.activeEdges.push_back((int)k);
      activeV[e.vi] = activeV[e.vj] = 1;
    }
  }
  // active vertices sorted by id; Hessian index = running count over non-fixed
  std::vector<std::pair<int, int>> byId;
  for (size_t v = 0; v < G.verts.size(); v++) {
    G.verts[v].hidx = -1;
    G.verts[v].bidx = -1;
    if (activeV[v]) byId.push_back({G.verts[v].id, (int)v});
  }
  std::sort(byId.begin(), byId.end());
  S.ivMap.clear();
  int off = 0;
  for (auto& pr : byId) {
    Vertex& v = G.verts[pr.second];
    if (!v.fixed) {
      v.bidx = (int)S.ivMap.size();
      v.hidx = off;
      off += v.dim;
      S.ivMap.push_back(pr.second);
    }
  }
  S.n = off;
  G.nfree = off;
  return !S.ivMap.empty();
}

// BlockSolver::buildStructure: upper-triangular block pattern + scalar CCS (fillCCS upper)
void buildStructure(Graph& G, System& S) {
  S.blockOf.clear();
  S.blocks.clear();
  auto addBlock = [&](int r, int c, int dr, int dc) {
    auto key = std::make_pair(r, c);
    if (S.blockOf.find(key) == S.blockOf.end()) {
      Block bl;
      bl.r = r; bl.c = c; bl.dr = dr; bl.dc = dc;
      std::memset(bl.v, 0, sizeof(bl.v));
      S.blockOf[key] = (int)S.blocks.size();
      S.blocks.push_back(bl);
    }
  };
  for (size_t i = 0; i < S.ivMap.size(); i++) {
    int d = G.verts[S.ivMap[i]].dim;
    addBlock((int)i, (int)i, d, d);
  }
  for (int k : S.activeEdges) {
    const Edge& e = G.edges[k];
    const Vertex& a = G.verts[e.vi];
    const Vertex& b = G.verts[e.vj];
    if (a.bidx >= 0 && b.bidx >= 0) {
      if (a.bidx <= b.bidx) addBlock(a.bidx, b.bidx, a.dim, b.dim);
      else addBlock(b.bidx, a.bidx, b.dim, a.dim);
    }
  }
  S.edgeBlk.assign(S.activeEdges.size() * 3, -1);
  for (size_t q = 0; q < S.activeEdges.size(); q++) {
    const Edge& e = G.edges[S.activeEdges[q]];
    const Vertex& a = G.verts[e.vi];
    const Vertex& b = G.verts[e.vj];
    if (a.bidx >= 0) S.edgeBlk[3 * q] = S.blockOf[{a.bidx, a.bidx}];
    if (b.bidx >= 0) S.edgeBlk[3 * q + 1] = S.blockOf[{b.bidx, b.bidx}];
    if (a.bidx >= 0 && b.bidx >= 0)
      S.edgeBlk[3 * q + 2] = S.blockOf[{std::min(a.bidx, b.bidx), std::max(a.bidx, b.bidx)}];
  }
  // scalar upper CSC
  int n = S.n;
  std::vector<int> boff(S.ivMap.size());
  for (size_t i = 0; i < S.ivMap.size(); i++) boff[i] = G.verts[S.ivMap[i]].hidx;
  std::vector<std::vector<std::pair<int, std::pair<int, int>>>> cols(n);  // row, (block, k)
  for (size_t bi = 0; bi < S.blocks.size(); bi++) {
    const Block& bl = S.blocks[bi];
    for (int r = 0; r < bl.dr; r++)
      for (int c = 0; c < bl.dc; c++) {
        int gr = boff[bl.r] + r, gc = boff[bl.c] + c;
        if (gr <= gc) cols[gc].push_back({gr, {(int)bi, r * bl.dc + c}});
      }
  }
  S.A.n = n;
  S.A.Ap.assign(n + 1, 0);
  S.A.Ai.clear();
  S.blockCsc.assign(S.blocks.size() * 9, -1);
  for (int j = 0; j < n; j++) {
    std::sort(cols[j].begin(), cols[j].end());
    for (auto& en : cols[j]) {
      S.blockCsc[en.second.first * 9 + en.second.second] = (int)S.A.Ai.size();
      S.A.Ai.push_back(en.first);
    }
    S.A.Ap[j + 1] = (int)S.A.Ai.size();
  }
  S.A.Ax.assign(S.A.Ai.size(), 0);
  S.b.assign(n, 0);
  S.x.assign(n, 0);
  G.nnzH = (long)S.A.Ai.size();
}

// BlockSolver::buildSystem: linearizeOplus + BaseBinaryEdge::constructQuadraticForm in edge order
void buildSystem(Graph& G, System& S) {
  for (auto& bl : S.blocks) std::memset(bl.v, 0, sizeof(bl.v));
  std::fill(S.b.begin(), S.b.end(), 0.0);
  for (size_t q = 0; q < S.activeEdges.size(); q++) {
    const Edge& e = G.edges[S.activeEdges[q]];
    const Vertex& from = G.verts[e.vi];
    const Vertex& to = G.verts[e.vj];
    int D = edgeDim(e), di = from.dim, dj = to.dim;
    double err[3], A[9], B[9];
    computeError(G, e, err);
    linearize(G, e, A, B);
    bool fromNotFixed = !from.fixed, toNotFixed = !to.fixed;
    double omega_r[3];
    for (int r = 0; r < D; r++) {
      double t = 0;
      for (int c = 0; c < D; c++) t += e.info[r * D + c] * err[c];
      omega_r[r] = -t;
    }
    double AtO[9], BtO[9];  // di x D, dj x D
    for (int r = 0; r < di; r++)
      for (int c = 0; c < D; c++) {
        double t = 0;
        for (int q = 0; q < D; q++) t += A[q * 3 + r] * e.info[q * D + c];
        AtO[r * 3 + c] = t;
      }
    for (int r = 0; r < dj; r++)
      for (int c = 0; c < D; c++) {
        double t = 0;
        for (int q = 0; q < D; q++) t += B[q * 3 + r] * e.info[q * D + c];
        BtO[r * 3 + c] = t;
      }
    if (fromNotFixed) {
      for (int r = 0; r < di; r++) {
        double t = 0;
        for (int q = 0; q < D; q++) t += A[q * 3 + r] * omega_r[q];
        S.b[from.hidx + r] += t;
      }
      Block& Hii = S.blocks[S.edgeBlk[3 * q]];
      for (int r = 0; r < di; r++)
        for (int c = 0; c < di; c++) {
          double t = 0;
          for (int q = 0; q < D; q++) t += AtO[r * 3 + q] * A[q * 3 + c];
          Hii.v[r * di + c] += t;
        }
      if (toNotFixed) {
        if (from.bidx <= to.bidx) {
          Block& Hij = S.blocks[S.edgeBlk[3 * q + 2]];
          for (int r = 0; r < di; r++)
            for (int c = 0; c < dj; c++) {
              double t = 0;
              for (int q = 0; q < D; q++) t += AtO[r * 3 + q] * B[q * 3 + c];
              Hij.v[r * dj + c] += t;
            }
        } else {  // _hessianRowMajor: write the transposed block
          Block& Hji = S.blocks[S.edgeBlk[3 * q + 2]];
          for (int r = 0; r < dj; r++)
            for (int c = 0; c < di; c++) {
              double t = 0;
              for (int q = 0; q < D; q++) t += B[q * 3 + r] * AtO[c * 3 + q];
              Hji.v[r * di + c] += t;
            }
        }
      }
    }
    if (toNotFixed) {
      for (int r = 0; r < dj; r++) {
        double t = 0;
        for (int q = 0; q < D; q++) t += B[q * 3 + r] * omega_r[q];
        S.b[to.hidx + r] += t;
      }
      Block& Hjj = S.blocks[S.edgeBlk[3 * q + 1]];
      for (int r = 0; r < dj; r++)
        for (int c = 0; c < dj; c++) {
          double t = 0;
          for (int q = 0; q < D; q++) t += BtO[r * 3 + q] * B[q * 3 + c];
          Hjj.v[r * dj + c] += t;
        }
    }
  }
  // fillCCS(upper)
  for (size_t bi = 0; bi < S.blocks.size(); bi++) {
    const Block& bl = S.blocks[bi];
    for (int k = 0; k < bl.dr * bl.dc; k++) {
      int q = S.blockCsc[bi * 9 + k];
      if (q >= 0) S.A.Ax[q] = bl.v[k];
    }
  }
}

double activeChi2(const Graph& G, const System& S) {
  double c = 0;
  for (int k : S.activeEdges) c += edgeChi2(G, G.edges[k]);
  return c;
}

}  // namespace

extern "C" {

void* orc_graph_create() { return new Graph(); }
void orc_graph_destroy(void* g) { delete static_cast<Graph*>(g); }

// slam.cpp:433-438 (VertexSE2, setEstimate(Vector3d) -> SE2(v): no angle normalisation)
int orc_graph_add_pose(void* g, int id, double x, double y, double th) {
  Graph& G = *static_cast<Graph*>(g);
  if (G.id2v.count(id)) return -1;  // g2o addVertex returns false on duplicate id
  Vertex v;
  v.id = id; v.dim = 3; v.est[0] = x; v.est[1] = y; v.est[2] = th; v.fixed = false; v.hidx = v.bidx = -1;
  G.id2v[id] = (int)G.verts.size();
  G.verts.push_back(v);
  return 0;
}
// slam.cpp:525-531
int orc_graph_add_landmark(void* g, int id, double x, double y) {
  Graph& G = *static_cast<Graph*>(g);
  if (G.id2v.count(id)) return -1;
  Vertex v;
  v.id = id; v.dim = 2; v.est[0] = x; v.est[1] = y; v.est[2] = 0; v.fixed = false; v.hidx = v.bidx = -1;
  G.id2v[id] = (int)G.verts.size();
  G.verts.push_back(v);
  return 0;
}
// slam.cpp:447-457 with an explicit measurement
int orc_graph_add_edge_se2(void* g, int idFrom, int idTo, const double* z3, const double* info9) {
  Graph& G = *static_cast<Graph*>(g);
  auto a = G.id2v.find(idFrom), b = G.id2v.find(idTo);
  if (a == G.id2v.end() || b == G.id2v.end()) return -1;
  Edge e;
  e.kind = 0; e.vi = a->second; e.vj = b->second;
  for (int k = 0; k < 3; k++) e.z[k] = z3[k];
  for (int k = 0; k < 9; k++) e.info[k] = info9[k];
  e.zinv = SE2(z3[0], z3[1], z3[2]).inverse();
  G.edges.push_back(e);
  return 0;
}
// slam.cpp:445-459: measurement = prevEstimate^-1 * SE2(pose)
int orc_graph_add_odometry(void* g, int idPrev, int idCur, const double* pose3, const double* info9) {
  Graph& G = *static_cast<Graph*>(g);
  auto a = G.id2v.find(idPrev);
  if (a == G.id2v.end()) return -1;
  const Vertex& pv = G.verts[a->second];
  SE2 prev(pv.est[0], pv.est[1], pv.est[2]);
  SE2 cur(pose3[0], pose3[1], pose3[2]);
  SE2 m = prev.inverse() * cur;
  double z[3] = {m.x, m.y, m.th};
  return orc_graph_add_edge_se2(g, idPrev, idCur, z, info9);
}
// the same measurement, returned (lets tests pin the GPU-side graph builder)
void orc_odometry_measurement(const double* prev3, const double* cur3, double* z3) {
  SE2 m = SE2(prev3[0], prev3[1], prev3[2]).inverse() * SE2(cur3[0], cur3[1], cur3[2]);
  z3[0] = m.x; z3[1] = m.y; z3[2] = m.th;
}
// slam.cpp:537-547
int orc_graph_add_edge_se2_xy(void* g, int poseId, int lmId, const double* z2, const double* info4) {
  Graph& G = *static_cast<Graph*>(g);
  auto a = G.id2v.find(poseId), b = G.id2v.find(lmId);
  if (a == G.id2v.end() || b == G.id2v.end()) return -1;
  Edge e;
  e.kind = 1; e.vi = a->second; e.vj = b->second;
  e.z[0] = z2[0]; e.z[1] = z2[1]; e.z[2] = 0;
  for (int k = 0; k < 9; k++) e.info[k] = 0;
  for (int k = 0; k < 4; k++) e.info[k] = info4[k];
  G.edges.push_back(e);
  return 0;
}
// slam.cpp:464-474
int orc_graph_set_fixed(void* g, int id, int fixed) {
  Graph& G = *static_cast<Graph*>(g);
  auto a = G.id2v.find(id);
  if (a == G.id2v.end()) return -1;
  G.verts[a->second].fixed = fixed != 0;
  return 0;
}
int orc_graph_get_vertex(void* g, int id, double* out3) {
  Graph& G = *static_cast<Graph*>(g);
  auto a = G.id2v.find(id);
  if (a == G.id2v.end()) return -1;
  const Vertex& v = G.verts[a->second];
  out3[0] = v.est[0]; out3[1] = v.est[1]; out3[2] = v.dim == 3 ? v.est[2] : 0;
  return v.dim;
}
int orc_graph_set_vertex(void* g, int id, const double* in3) {
  Graph& G = *static_cast<Graph*>(g);
  auto a = G.id2v.find(id);
  if (a == G.id2v.end()) return -1;
  Vertex& v = G.verts[a->second];
  for (int k = 0; k < v.dim; k++) v.est[k] = in3[k];
  return 0;
}
// bulk read-back in insertion order: out[3*k..] (landmarks leave slot 2 = 0)
int orc_graph_get_all(void* g, int* ids, double* out3, int cap) {
  Graph& G = *static_cast<Graph*>(g);
  int n = (int)G.verts.size();
  for (int k = 0; k < n && k < cap; k++) {
    if (ids) ids[k] = G.verts[k].id;
    out3[3 * k] = G.verts[k].est[0];
    out3[3 * k + 1] = G.verts[k].est[1];
    out3[3 * k + 2] = G.verts[k].dim == 3 ? G.verts[k].est[2] : 0;
  }
  return n;
}
int orc_graph_num_vertices(void* g) { return (int)static_cast<Graph*>(g)->verts.size(); }
int orc_graph_num_edges(void* g) { return (int)static_cast<Graph*>(g)->edges.size(); }

double orc_graph_chi2(void* g) {
  Graph& G = *static_cast<Graph*>(g);
  System S;
  initializeOptimization(G, S);
  return activeChi2(G, S);
}

// initializeOptimization(); optimize(iters) with verbose chi2 (slam.cpp:63, 480-481).
// chi2[k] = active chi2 after iteration k's update, k < return value.
// Returns: -1 nothing to optimise, 0 factorisation failure, else number of iterations.
int orc_graph_optimize(void* g, int iters, double* chi2) {
  Graph& G = *static_cast<Graph*>(g);
  System S;
  double t0 = now_s();
  G.t_struct = G.t_linearize = G.t_analyze = G.t_factor = G.t_solve = G.t_update = 0;
  bool any = initializeOptimization(G, S);
  G.t_init = now_s() - t0;
  if (!any) return -1;
#ifdef ORACLE_USE_EIGEN
  typedef Eigen::SparseMatrix<double, Eigen::ColMajor> SpMat;
  Eigen::SimplicialLDLT<SpMat, Eigen::Upper> chol;  // g2o LinearSolverEigen, blockOrdering=false
  SpMat M;
#else
  SimplicialLDLT chol;
#endif
  int done = 0;
  bool ok = true;
  for (int it = 0; it < iters && ok; it++) {
    double t1 = now_s();
    if (it == 0) {
      buildStructure(G, S);
      G.t_struct = now_s() - t1;
      t1 = now_s();
    }
    buildSystem(G, S);
    G.t_linearize += now_s() - t1;
    t1 = now_s();
#ifdef ORACLE_USE_EIGEN
    if (it == 0) {
      M = Eigen::Map<SpMat>(S.n, S.n, (int)S.A.Ai.size(), S.A.Ap.data(), S.A.Ai.data(), S.A.Ax.data());
      chol.analyzePattern(M);
      G.t_analyze = now_s() - t1;
      t1 = now_s();
    } else {
      std::memcpy(M.valuePtr(), S.A.Ax.data(), sizeof(double) * S.A.Ax.size());
    }
    chol.factorize(M);
    ok = chol.info() == Eigen::Success;
    G.t_factor += now_s() - t1;
    t1 = now_s();
    if (ok) {
      Eigen::Map<Eigen::VectorXd> bb(S.b.data(), S.n), xx(S.x.data(), S.n);
      xx = chol.solve(bb);
      if (it == 0) {
        G.nnzL = 0; G.flopsL = 0;
        SpMat Lm = chol.matrixL();
        for (int k = 0; k < Lm.outerSize(); k++) {
          double c = Lm.outerIndexPtr()[k + 1] - Lm.outerIndexPtr()[k] - 1;  // unit diagonal stored
          if (c < 0) c = 0;
          G.nnzL += (long)c;
          G.flopsL += c * c + 3 * c;
        }
      }
    }
#else
    if (it == 0) {
      chol.analyze(S.A);
      G.t_analyze = now_s() - t1;
      t1 = now_s();
      G.nnzL = chol.Lp[S.n];
      G.flopsL = chol.flops;
    }
    ok = chol.factorize(S.A);
    G.t_factor += now_s() - t1;
    t1 = now_s();
    if (ok) chol.solve(S.b.data(), S.x.data());
#endif
    G.t_solve += now_s() - t1;
    if (!ok) break;
    t1 = now_s();
    // SparseOptimizer::update
    for (size_t i = 0; i < S.ivMap.size(); i++) {
      Vertex& v = G.verts[S.ivMap[i]];
      oplus(v, S.x.data() + v.hidx);
    }
    if (chi2) chi2[it] = activeChi2(G, S);  // verbose: computeActiveErrors + activeRobustChi2
    G.t_update += now_s() - t1;
    done++;
  }
  G.t_total = now_s() - t0;
  if (!ok) return 0;
  return done;
}

// timings / sizes of the last optimize: [init, struct, linearize, analyze, factor, solve, update,
// total, nfree, nnzH, nnzL, flopsL]
void orc_graph_stats(void* g, double* out12) {
  Graph& G = *static_cast<Graph*>(g);
  out12[0] = G.t_init; out12[1] = G.t_struct; out12[2] = G.t_linearize; out12[3] = G.t_analyze;
  out12[4] = G.t_factor; out12[5] = G.t_solve; out12[6] = G.t_update; out12[7] = G.t_total;
  out12[8] = G.nfree; out12[9] = (double)G.nnzH; out12[10] = (double)G.nnzL; out12[11] = G.flopsL;
}

// One linearisation at the current estimate, exported for kernel-level parity tests:
// the scalar upper-triangular CSC of H, b (= -J^T Omega e), the Hessian offset of every vertex in
// insertion order (-1 = fixed/inactive) and chi2.  Call with Ai == NULL to query sizes
// (returns nnz; *n_out = scalar dimension).
long orc_graph_build_system(void* g, int* n_out, int* hidx_out, int* Ap, int* Ai, double* Ax,
                            double* b, double* chi2) {
  Graph& G = *static_cast<Graph*>(g);
  System S;
  if (!initializeOptimization(G, S)) {
    *n_out = 0;
    return 0;
  }
  buildStructure(G, S);
  *n_out = S.n;
  if (hidx_out)
    for (size_t v = 0; v < G.verts.size(); v++) hidx_out[v] = G.verts[v].hidx;
  if (!Ai) return (long)S.A.Ai.size();
  buildSystem(G, S);
  std::memcpy(Ap, S.A.Ap.data(), sizeof(int) * (S.n + 1));
  std::memcpy(Ai, S.A.Ai.data(), sizeof(int) * S.A.Ai.size());
  std::memcpy(Ax, S.A.Ax.data(), sizeof(double) * S.A.Ax.size());
  std::memcpy(b, S.b.data(), sizeof(double) * S.n);
  if (chi2) *chi2 = activeChi2(G, S);
  return (long)S.A.Ai.size();
}

// per-edge error and Jacobians (row-major, leading dimension 3) for finite-difference tests
int orc_graph_edge_linearization(void* g, int edge, double* err3, double* Ji9, double* Jj9) {
  Graph& G = *static_cast<Graph*>(g);
  if (edge < 0 || edge >= (int)G.edges.size()) return -1;
  computeError(G, G.edges[edge], err3);
  linearize(G, G.edges[edge], Ji9, Jj9);
  return edgeDim(G.edges[edge]);
}

double orc_normalize_theta(double t) { return normalize_theta(t); }

int orc_uses_eigen() {
#ifdef ORACLE_USE_EIGEN
  return 1;
#else
  return 0;
#endif
}

}  // extern "C"

// =============================================================================================
// The Slam back half as one object: performSLAM (slam.cpp:298-338) and everything below it, with
// the pose handed in explicitly instead of read from m_odometryData (the front half -- message
// handling, frame gathering, keyframing -- is out of scope and wall-clock dependent).
// =============================================================================================
namespace {
struct OracleSlam {
  void* graph = orc_graph_create();
  std::vector<double> map_x, map_y;
  std::vector<int> map_type;
  std::vector<double> poses;                       // m_poses, 3 per pose
  std::vector<std::vector<int>> connectivity;      // m_connectivityGraph
  double newConeThreshold = 1, coneMappingThreshold = 67;  // slam.hpp:113,116
  unsigned currentConeIndex = 0;                   // slam.hpp:117
  int poseId = 1000;                               // slam.hpp:118
  int loopClosing = 0, loopClosingComplete = 0;    // slam.hpp:123-124
  double sendPose[3] = {0, 0, 0};
  int optimizeCalls = 0;
  std::vector<double> chi2Log;                     // chi2 of every GN iteration ever run
  int lastIterations = 0;
  // SURVEY 8(f) rank 3, OPT-IN (default off = the reference's behaviour, bug included): see
  // orc_slam_set_localizer_repair below.
  int localizerRepair = 0, localizerWindow = 10;
  int landmarksFrozen = 0, nextPoseToFix = 1000;
  ~OracleSlam() { orc_graph_destroy(graph); }

  void addConeMeasurement(int coneId, const double m3[3]) {  // 537-550
    double s[3];
    spherical2Cartesian(m3[0], m3[1], m3[2], s);
    double info[4] = {0.01, 0, 0, 0.01};
    orc_graph_add_edge_se2_xy(graph, poseId - 1, coneId, s, info);
    connectivity[poseId - 1001].push_back(coneId);
  }
  void optimizeGraph() {  // 461-484
    orc_graph_set_fixed(graph, 1000, 1);
    orc_graph_set_fixed(graph, 1001, 1);
    orc_graph_set_fixed(graph, 0, 1);
    orc_graph_set_fixed(graph, 1, 1);
    double chi2[10];
    lastIterations = orc_graph_optimize(graph, 10, chi2);
    for (int k = 0; k < lastIterations; k++) chi2Log.push_back(chi2[k]);
    optimizeCalls++;
  }
  // Repaired localiser only: the optimise slam.cpp:403 leaves commented out, as a sliding window.
  // The map is frozen after loop closure (every landmark fixed, once), poses older than the last
  // `localizerWindow` ones are fixed for good as the window moves on, the gauge is the reference's.
  void optimizeWindow() {
    if (!landmarksFrozen) {
      for (size_t j = 0; j < map_x.size(); j++) orc_graph_set_fixed(graph, (int)j, 1);
      landmarksFrozen = 1;
    }
    for (; nextPoseToFix < poseId - localizerWindow; nextPoseToFix++) orc_graph_set_fixed(graph, nextPoseToFix, 1);
    optimizeGraph();
  }
  void updateMap() {  // 713-732
    for (size_t j = 0; j < map_x.size(); j++) {
      double e[3];
      orc_graph_get_vertex(graph, (int)j, e);
      map_x[j] = e[0];
      map_y[j] = e[1];
    }
  }
};
}  // namespace

extern "C" {

void* orc_slam_create(double sameConeThreshold, double coneMappingThreshold) {
  OracleSlam* s = new OracleSlam();
  s->newConeThreshold = sameConeThreshold;
  s->coneMappingThreshold = coneMappingThreshold;
  return s;
}
void orc_slam_destroy(void* p) { delete static_cast<OracleSlam*>(p); }
// SURVEY 8(f) rank 3 (behaviour change, hence opt-in): localiser frames add pose -> cone edges whose
// measurement is the observation instead of the pose slam.cpp:373 passes, and run the optimise that
// slam.cpp:403 comments out over the last `window` poses against the frozen map.
void orc_slam_set_localizer_repair(void* p, int on, int window) {
  OracleSlam& S = *static_cast<OracleSlam*>(p);
  S.localizerRepair = on != 0;
  S.localizerWindow = window > 0 ? window : 1;
}

// performSLAM.  idx/status (N each) receive the per-observation association records of this
// frame (mapping phase: as orc_assoc_map_frame; localisation phase: status 0/2).
// Returns 0 = mapping frame, 1 = mapping frame that closed the loop (optimised), 2 = localiser
// frame, -1 = rejected by the 200 m gate (300-303).
int orc_slam_perform(void* p, const double* cones4xN, int N, const double* pose_in, double yawRate,
                     double timeElapsed, int* idx, int* status) {
  OracleSlam& S = *static_cast<OracleSlam*>(p);
  if (std::fabs(pose_in[0]) > 200 || std::fabs(pose_in[1]) > 200) return -1;
  double pose[3] = {pose_in[0], pose_in[1], pose_in[2]};
  if (timeElapsed > 0 && timeElapsed < 1) pose[2] = pose[2] - yawRate * timeElapsed;  // 315-317
  S.poses.insert(S.poses.end(), pose, pose + 3);
  // addPoseToGraph 433-443
  orc_graph_add_pose(S.graph, S.poseId, pose[0], pose[1], pose[2]);
  if (S.poseId > 1000) {
    double info[9] = {5, 0, 0, 0, 5, 0, 0, 0, 5};
    orc_graph_add_odometry(S.graph, S.poseId - 1, S.poseId, pose, info);
  }
  S.connectivity.push_back(std::vector<int>());
  S.poseId++;
  int ret = 0;
  for (int i = 0; i < N; i++) { idx[i] = -1; status[i] = 2; }
  if (!S.loopClosingComplete) {  // addConesToMap 552-635
    int M = (int)S.map_x.size();
    int cap = M + N + 1;
    S.map_x.resize(cap); S.map_y.resize(cap); S.map_type.resize(cap);
    std::vector<double> z(2 * (size_t)N + 2), g(3 * (size_t)N + 3);
    int first = 0, lcObs = -1;
    orc_assoc_map_frame(cones4xN, N, pose, S.newConeThreshold, S.coneMappingThreshold,
                        S.map_x.data(), S.map_y.data(), S.map_type.data(), &M, cap,
                        &S.currentConeIndex, &S.loopClosing, idx, status, z.data(), g.data(),
                        &first, &lcObs);
    S.map_x.resize(M); S.map_y.resize(M); S.map_type.resize(M);
    if (first) {  // 554-567
      orc_graph_add_landmark(S.graph, 0, S.map_x[0], S.map_y[0]);
      S.addConeMeasurement(0, cones4xN);
    }
    for (int i = 0; i < N; i++) {
      const double* col = cones4xN + 4 * (size_t)i;
      if (status[i] == 0) S.addConeMeasurement(idx[i], col);                     // 591
      else if (status[i] == 1) {                                                 // 610-619
        orc_graph_add_landmark(S.graph, idx[i], g[3 * i], g[3 * i + 1]);
        S.addConeMeasurement(idx[i], col);
      }
    }
    if (S.loopClosing) {  // 625-633: once per observation from the triggering one on
      int from = lcObs >= 0 ? lcObs : 0;
      for (int i = from; i < N; i++) {
        S.optimizeGraph();
        S.updateMap();
        S.loopClosingComplete = 1;
      }
      ret = 1;
    }
  }
  if (S.loopClosingComplete && N > 1) {  // localizer 340-414
    std::vector<double> g(3 * (size_t)N);
    std::vector<int> lidx(N);
    int reobs = 0, send = 0;
    orc_assoc_localize_frame(cones4xN, N, pose, S.newConeThreshold, S.map_x.data(), S.map_y.data(),
                             S.map_type.data(), (int)S.map_x.size(), &S.currentConeIndex,
                             lidx.data(), g.data(), &reobs, &send);
    if (S.localizerRepair) {
      // opt-in repair: the edge carries the OBSERVATION (what 373 meant), then the window optimise
      for (int i = 0; i < N; i++)
        if (lidx[i] >= 0) S.addConeMeasurement(lidx[i], cones4xN + 4 * (size_t)i);
      if (reobs > 0) S.optimizeWindow();
    } else {
      for (int i = 0; i < N; i++)
        if (lidx[i] >= 0) S.addConeMeasurement(lidx[i], pose);  // 373: passes the POSE (sic)
    }
    if (ret == 0) {
      for (int i = 0; i < N; i++) { idx[i] = lidx[i]; status[i] = lidx[i] >= 0 ? 0 : 2; }
      ret = 2;
    }
    orc_graph_get_vertex(S.graph, S.poseId - 1, S.sendPose);  // updatePoseFromGraph 416-422
  }
  return ret;
}

int orc_slam_map_size(void* p) { return (int)static_cast<OracleSlam*>(p)->map_x.size(); }
void orc_slam_get_map(void* p, double* x, double* y, int* type) {
  OracleSlam& S = *static_cast<OracleSlam*>(p);
  for (size_t j = 0; j < S.map_x.size(); j++) { x[j] = S.map_x[j]; y[j] = S.map_y[j]; type[j] = S.map_type[j]; }
}
// state: [currentConeIndex, poseId, loopClosing, loopClosingComplete, optimizeCalls, lastIterations,
//         nChi2, nEdges]
void orc_slam_state(void* p, int* out8) {
  OracleSlam& S = *static_cast<OracleSlam*>(p);
  out8[0] = (int)S.currentConeIndex; out8[1] = S.poseId; out8[2] = S.loopClosing;
  out8[3] = S.loopClosingComplete; out8[4] = S.optimizeCalls; out8[5] = S.lastIterations;
  out8[6] = (int)S.chi2Log.size(); out8[7] = orc_graph_num_edges(S.graph);
}
// row k of the connectivity graph (cone ids addConeMeasurement recorded for pose 1000 + k, slam.cpp:549);
// returns its length (ids are written up to cap), -1 if there is no such row
int orc_slam_connectivity_row(void* p, int k, int* out, int cap) {
  OracleSlam& S = *static_cast<OracleSlam*>(p);
  if (k < 0 || k >= (int)S.connectivity.size()) return -1;
  const std::vector<int>& r = S.connectivity[k];
  for (size_t i = 0; i < r.size() && (int)i < cap; i++) out[i] = r[i];
  return (int)r.size();
}
int orc_slam_num_poses(void* p) { return (int)(static_cast<OracleSlam*>(p)->poses.size() / 3); }
void orc_slam_get_poses(void* p, double* out3n) {
  OracleSlam& S = *static_cast<OracleSlam*>(p);
  for (size_t k = 0; k < S.poses.size(); k++) out3n[k] = S.poses[k];
}
void orc_slam_chi2_log(void* p, double* out) {
  OracleSlam& S = *static_cast<OracleSlam*>(p);
  for (size_t k = 0; k < S.chi2Log.size(); k++) out[k] = S.chi2Log[k];
}
int orc_slam_get_pose(void* p, int id, double* out3) {
  return orc_graph_get_vertex(static_cast<OracleSlam*>(p)->graph, id, out3);
}
void orc_slam_send_pose(void* p, double* out3) {
  OracleSlam& S = *static_cast<OracleSlam*>(p);
  out3[0] = S.sendPose[0]; out3[1] = S.sendPose[1]; out3[2] = S.sendPose[2];
}
void* orc_slam_graph(void* p) { return static_cast<OracleSlam*>(p)->graph; }

}  // extern "C"

// Bulk loader (same insertion order as performSLAM produces: all landmark and pose vertices, then
// per pose its incoming odometry edge(s) followed by its cone edges), so large graphs load fast.
extern "C" int orc_graph_load(void* g, int P, const int* pose_ids, const double* pose_est3, int L,
                              const int* lm_ids, const double* lm_est2, int Eo, const int* eo_from,
                              const int* eo_to, const double* eo_z3, const double* eo_info9, int El,
                              const int* el_pose, const int* el_lm, const double* el_z2,
                              const double* el_info4, int nfixed, const int* fixed_ids) {
  for (int l = 0; l < L; l++)
    if (orc_graph_add_landmark(g, lm_ids[l], lm_est2[2 * l], lm_est2[2 * l + 1])) return -1;
  for (int p = 0; p < P; p++)
    if (orc_graph_add_pose(g, pose_ids[p], pose_est3[3 * p], pose_est3[3 * p + 1], pose_est3[3 * p + 2])) return -1;
  std::map<int, int> pos;
  for (int p = 0; p < P; p++) pos[pose_ids[p]] = p;
  std::vector<std::vector<int>> eoBy(P), elBy(P);
  for (int e = 0; e < Eo; e++) {
    auto it = pos.find(eo_to[e]);
    if (it == pos.end()) return -1;
    eoBy[it->second].push_back(e);
  }
  for (int e = 0; e < El; e++) {
    auto it = pos.find(el_pose[e]);
    if (it == pos.end()) return -1;
    elBy[it->second].push_back(e);
  }
  for (int p = 0; p < P; p++) {
    for (int e : eoBy[p])
      if (orc_graph_add_edge_se2(g, eo_from[e], eo_to[e], eo_z3 + 3 * (size_t)e, eo_info9 + 9 * (size_t)e)) return -1;
    for (int e : elBy[p])
      if (orc_graph_add_edge_se2_xy(g, el_pose[e], el_lm[e], el_z2 + 2 * (size_t)e, el_info4 + 4 * (size_t)e)) return -1;
  }
  for (int k = 0; k < nfixed; k++)
    if (orc_graph_set_fixed(g, fixed_ids[k], 1)) return -1;
  return 0;
}

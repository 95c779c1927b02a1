// slam.cpp -- back half of class Slam over the C ABI (see slam.hpp).  Each method cites the
// reference lines whose behaviour it reproduces; none of the arithmetic happens here -- the
// conversion, gating, linearisation, assembly and solve run in the CUDA library.
#include "slam.hpp"

#include <algorithm>
#include <cmath>
#include <iostream>
#include <stdexcept>

using slamtypes::MatrixXd;
using slamtypes::Vector3d;

namespace {
void check(slam_b200_ctx* ctx, int rc, const char* what) {
  if (rc < 0) {
    std::string msg = std::string(what) + " failed (" + std::to_string(rc) + "): " +
                      (ctx ? slam_b200_last_error(ctx) : "no context");
    throw std::runtime_error(msg);
  }
}
}  // namespace

Slam::Slam(std::map<std::string, std::string> commandlineArguments)
    : m_odometryData(), m_gpsReference(), m_map(), m_sendPose() {
  setUp(commandlineArguments);
  int device = 0;
  auto it = commandlineArguments.find("cudaDevice");
  if (it != commandlineArguments.end()) device = std::stoi(it->second);
  it = commandlineArguments.find("localizerWindow");
  if (it != commandlineArguments.end()) m_localizerWindow = std::max(1, std::stoi(it->second));
  it = commandlineArguments.find("localizerRepair");
  if (it != commandlineArguments.end()) m_localizerRepair = std::stoi(it->second) != 0;
  // setupOptimizer (slam.cpp:53-65): the GN / block solver / linear solver stack lives in the backend
  int rc = slam_b200_create(device, nullptr, &m_ctx);
  if (rc != 0) throw std::runtime_error("slam_b200_create failed: no CUDA device (no CPU fallback)");
  setupOptimizer();
  // first-use costs (kernel loading, attributes, device arrays of a lap-sized graph, frame mailbox) are paid
  // here, like the reference pays its solver set-up in its constructor, not in the loop-closing frame
  it = commandlineArguments.find("warmupPoses");
  const int warmupPoses = it != commandlineArguments.end() ? std::stoi(it->second) : 1024;
  if (warmupPoses > 0) check(m_ctx, slam_b200_warmup(m_ctx, warmupPoses, std::max(8, warmupPoses / 3)), "warmup");
  m_odometryData = Vector3d(0, 0, 0);
  m_sendPose = Vector3d(0, 0, 0);
}

Slam::~Slam() { tearDown(); }

void Slam::setupOptimizer() {
  slam_b200_graph_clear(m_ctx);
  slam_b200_map_clear(m_ctx);
}

void Slam::tearDown() {
  if (m_ctx) slam_b200_destroy(m_ctx);
  m_ctx = nullptr;
}

// slam.cpp:736-756
void Slam::setUp(std::map<std::string, std::string> configuration) {
  m_timeDiffMilliseconds = static_cast<uint32_t>(std::stoi(configuration["gatheringTimeMs"]));
  m_newConeThreshold = static_cast<double>(std::stod(configuration["sameConeThreshold"]));
  m_gpsReference[0] = static_cast<double>(std::stod(configuration["refLatitude"]));
  m_gpsReference[1] = static_cast<double>(std::stod(configuration["refLongitude"]));
  m_timeBetweenKeyframes = static_cast<double>(std::stod(configuration["timeBetweenKeyframes"]));
  m_coneMappingThreshold = static_cast<double>(std::stod(configuration["coneMappingThreshold"]));
  m_conesPerPacket = static_cast<int>(std::stoi(configuration["conesPerPacket"]));
  m_senderStamp = static_cast<int>(std::stoi(configuration["id"]));
}

void Slam::setOdometry(double x, double y, double heading) {
  std::lock_guard<std::mutex> lockSensor(m_sensorMutex);
  m_odometryData = Vector3d(x, y, heading);
}

void Slam::setYawRate(float yawRate, double secondsSinceYaw) {
  std::lock_guard<std::mutex> lockYaw(m_yawMutex);
  m_yawRate = yawRate;
  m_yawElapsed = secondsSinceYaw;
}

// slam.cpp:298-338
void Slam::performSLAM(MatrixXd cones) {
  m_lastFrameKind = -1;
  m_lastIdx.assign((size_t)cones.cols(), -1);
  m_lastStatus.assign((size_t)cones.cols(), SLAM_B200_ASSOC_NONE);
  if (std::fabs(m_odometryData(0)) > 200 || std::fabs(m_odometryData(1)) > 200) return;  // 300-303
  Vector3d pose;
  {
    std::lock_guard<std::mutex> lockSensor(m_sensorMutex);
    pose = m_odometryData;
    {
      std::lock_guard<std::mutex> lockYaw(m_yawMutex);
      double timeElapsed = m_yawElapsed;
      if (timeElapsed > 0 && timeElapsed < 1) pose(2) = pose(2) - static_cast<double>(m_yawRate) * (timeElapsed);  // 315-317
    }
    m_poses.push_back(pose);
  }
  {
    std::lock_guard<std::mutex> lockOptimizer(m_optimizerMutex);
    addPoseToGraph(pose);
  }
  m_lastFrameKind = 0;
  if (!m_loopClosingComplete) addConesToMap(cones, pose);               // 328-331
  if (m_loopClosingComplete && cones.cols() > 1) localizer(pose, cones);  // 332-334
}

// The host mirror and the device map grow in lock step; an association record that points outside the mirror
// means they have come apart (a backend fault) -- fail loudly instead of reading past m_map.
void Slam::checkMapIndex(int32_t j) const {
  if (j < 0 || (size_t)j >= m_map.size())
    throw std::runtime_error("association index " + std::to_string(j) + " outside the map of " +
                             std::to_string(m_map.size()) + " cones");
}

// slam.cpp:433-443
void Slam::addPoseToGraph(Vector3d pose) {
  check(m_ctx, slam_b200_graph_add_pose(m_ctx, m_poseId, pose(0), pose(1), pose(2)), "graph_add_pose");
  addOdometryMeasurement(pose);
  {
    // drawGraph() copies m_connectivityGraph under map + sensor mutex (slam.cpp:780-784) while the reference
    // grows it here under the optimizer mutex only: a viewer-thread read racing a reallocation.  The sensor
    // mutex is a leaf lock everywhere else, so taking it for the push_back cannot invert an order.
    std::lock_guard<std::mutex> lockSensor(m_sensorMutex);
    m_connectivityGraph.push_back(std::vector<int>());
  }
  m_poseId++;
}

// slam.cpp:445-459: EdgeSE2 prev -> cur, measurement prevEstimate^-1 * SE2(pose), information 5 I
void Slam::addOdometryMeasurement(Vector3d pose) {
  if (m_poseId > 1000) {
    const double info[9] = {5, 0, 0, 0, 5, 0, 0, 0, 5};
    const double p[3] = {pose(0), pose(1), pose(2)};
    check(m_ctx, slam_b200_graph_add_odometry(m_ctx, m_poseId - 1, m_poseId, p, info), "graph_add_odometry");
  }
}

// slam.cpp:525-535
void Slam::addConeToGraph(Cone cone, const double xy[2]) {
  check(m_ctx, slam_b200_graph_add_landmark(m_ctx, cone.getId(), cone.getX(), cone.getY()), "graph_add_landmark");
  addConeMeasurement(cone, xy);
}

// slam.cpp:537-550: EdgeSE2PointXY pose(m_poseId-1) -> cone, information 0.01 I.  The vehicle-frame
// xy measurement (Spherical2Cartesian of the observation, 539) comes from the association kernel.
void Slam::addConeMeasurement(Cone cone, const double xy[2]) {
  const double info[4] = {0.01, 0, 0, 0.01};
  check(m_ctx, slam_b200_graph_add_edge_se2_xy(m_ctx, m_poseId - 1, cone.getId(), xy, info), "graph_add_edge_se2_xy");
  m_connectivityGraph[m_poseId - 1001].push_back(cone.getId());
}

// slam.cpp:552-635
void Slam::addConesToMap(MatrixXd cones, Vector3d pose) {
  std::lock_guard<std::mutex> lockMap(m_mapMutex);
  const int n = (int)cones.cols();
  if (n == 0) return;
  std::vector<double> z(2 * (size_t)n), g(3 * (size_t)n);
  int32_t lc = m_loopClosing ? 1 : 0, first = 0, lcObs = -1;
  const double p[3] = {pose(0), pose(1), pose(2)};
  int rc = slam_b200_assoc_map_frame(m_ctx, cones.data(), n, p, m_newConeThreshold, m_coneMappingThreshold,
                                     &m_currentConeIndex, &lc, m_lastIdx.data(), m_lastStatus.data(), z.data(),
                                     g.data(), &first, &lcObs);
  check(m_ctx, rc, "assoc_map_frame");
  {
    std::lock_guard<std::mutex> lockOptimizer(m_optimizerMutex);
    if (first) {  // 554-567: cone 0 from column 0, its vertex and first edge
      Cone cone(g[0], g[1], (int)g[2], 0);
      m_map.push_back(cone);
      addConeToGraph(cone, &z[0]);
    }
    for (int i = 0; i < n; i++) {
      if (m_lastStatus[i] == SLAM_B200_ASSOC_MATCHED) {            // 584-592
        checkMapIndex(m_lastIdx[i]);
        addConeMeasurement(m_map[m_lastIdx[i]], &z[2 * (size_t)i]);
      } else if (m_lastStatus[i] == SLAM_B200_ASSOC_NEW) {         // 608-619
        Cone cone(g[3 * (size_t)i], g[3 * (size_t)i + 1], (int)g[3 * (size_t)i + 2], (int)m_map.size());
        m_map.push_back(cone);
        addConeToGraph(cone, &z[2 * (size_t)i]);
      }
    }
    m_loopClosing = lc != 0;
    if (m_loopClosing) {  // 625-633: one optimise + map update per column from the closing one on
      const int from = lcObs >= 0 ? lcObs : 0;
      for (int i = from; i < n; i++) {
        optimizeGraph();
        updateMap();
        m_loopClosingComplete = true;
      }
      m_lastFrameKind = 1;
    }
  }
}

// slam.cpp:461-484: fix the gauge, initializeOptimization(), optimize(10)
void Slam::optimizeGraph() {
  slam_b200_graph_set_fixed(m_ctx, 1000, 1);
  slam_b200_graph_set_fixed(m_ctx, 1001, 1);
  slam_b200_graph_set_fixed(m_ctx, 0, 1);
  slam_b200_graph_set_fixed(m_ctx, 1, 1);
  double chi2[10];
  int it = slam_b200_graph_optimize(m_ctx, 10, chi2);
  if (it < -1) check(m_ctx, it, "graph_optimize");
  m_lastIterations = it;
  for (int k = 0; k < it; k++) m_chi2Log.push_back(chi2[k]);
  m_optimizeCalls++;
}

void Slam::setLocalizerRepair(bool on, int window) {
  std::lock_guard<std::mutex> lockOptimizer(m_optimizerMutex);
  m_localizerRepair = on;
  m_localizerWindow = std::max(1, window);
}

// Repaired localiser only (opt-in): the optimise of slam.cpp:403 over the last m_localizerWindow poses
// against the frozen map.  Landmarks are fixed once, poses leave the window for good as it moves on.
void Slam::optimizeWindow() {
  if (!m_landmarksFrozen) {
    for (size_t j = 0; j < m_map.size(); j++) slam_b200_graph_set_fixed(m_ctx, (int)j, 1);
    m_landmarksFrozen = true;
  }
  for (; m_nextPoseToFix < m_poseId - m_localizerWindow; m_nextPoseToFix++)
    slam_b200_graph_set_fixed(m_ctx, m_nextPoseToFix, 1);
  optimizeGraph();
}

// slam.cpp:713-732: landmark estimates -> map.  The device map takes them on the device, behind the optimise on the
// back end's stream; the host-side m_map follows from the pinned mirror the back end refreshes asynchronously, and
// only when somebody reads it (the frame loop does not wait for a device -> host copy).
void Slam::updateMap() {
  check(m_ctx, slam_b200_map_update_from_graph(m_ctx), "map_update_from_graph");
  m_mapMirrorPending = true;
}

// called with m_mapMutex held, by everything that hands m_map's coordinates out
void Slam::refreshMapFromMirror() {
  if (!m_mapMirrorPending) return;
  const double *x = nullptr, *y = nullptr;
  const int n = slam_b200_map_mirror(m_ctx, &x, &y, nullptr);
  check(m_ctx, n, "map_mirror");
  for (size_t j = 0; j < m_map.size() && j < (size_t)n; j++) {
    m_map[j].setX(x[j]);
    m_map[j].setY(y[j]);
  }
  m_mapMirrorPending = false;
}

// slam.cpp:416-422
Vector3d Slam::updatePoseFromGraph() {
  double e[3] = {0, 0, 0};
  slam_b200_graph_get_vertex(m_ctx, m_poseId - 1, e);
  return Vector3d(e[0], e[1], e[2]);
}

// slam.cpp:340-414
void Slam::localizer(Vector3d pose, MatrixXd cones) {
  const int n = (int)cones.cols();
  const double p[3] = {pose(0), pose(1), pose(2)};
  std::vector<int32_t> idx((size_t)n, -1);
  int32_t reobs = 0, send = 0;
  {
    std::lock_guard<std::mutex> lockMap(m_mapMutex);
    int rc = slam_b200_assoc_localize_frame(m_ctx, cones.data(), n, p, m_newConeThreshold, &m_currentConeIndex,
                                            idx.data(), nullptr, &reobs, &send);
    check(m_ctx, rc, "assoc_localize_frame");
    std::lock_guard<std::mutex> lockOptimizer(m_optimizerMutex);
    if (m_localizerRepair) {
      if (reobs > 0) {
        // opt-in repair: the edge carries Spherical2Cartesian(observation), what 373 meant
        std::vector<double> local(3 * (size_t)n);
        check(m_ctx, slam_b200_cones_to_global(m_ctx, cones.data(), n, p, nullptr, local.data()), "cones_to_global");
        for (int i = 0; i < n; i++)
          if (idx[i] >= 0) {
            checkMapIndex(idx[i]);
            addConeMeasurement(m_map[idx[i]], &local[3 * (size_t)i]);
          }
        optimizeWindow();
      }
    } else if (reobs > 0) {
      // 373: addConeMeasurement(m_map[j], pose) -- the reference hands the POSE where a measurement
      // (azimuth, zenith, range) is expected; reproduced: Spherical2Cartesian(pose) is the edge's z.
      const double asObs[4] = {pose(0), pose(1), pose(2), 0};
      const double zero[3] = {0, 0, 0};
      double local[3];
      check(m_ctx, slam_b200_cones_to_global(m_ctx, asObs, 1, zero, nullptr, local), "cones_to_global");
      for (int i = 0; i < n; i++)
        if (idx[i] >= 0) {
          checkMapIndex(idx[i]);
          addConeMeasurement(m_map[idx[i]], local);
        }
    }
    m_sendConeData = send != 0;  // 385
  }
  if (m_lastFrameKind == 0) {
    m_lastFrameKind = 2;
    for (int i = 0; i < n; i++) {
      m_lastIdx[i] = idx[i];
      m_lastStatus[i] = idx[i] >= 0 ? SLAM_B200_ASSOC_MATCHED : SLAM_B200_ASSOC_NONE;
    }
  }
  // The reference keeps m_optimizerMutex from 402 to the end of the function, i.e. through sendCones()
  // (optimizer -> send -> map, 402 -> 660 -> 663), while addConesToMap takes map -> optimizer (553 -> 586):
  // an ABBA pair between two detached frame threads.  Here the lock covers the graph
  // read only; nothing the senders touch is guarded by it.
  Vector3d updatedPoseVectorGraph;
  {
    std::lock_guard<std::mutex> lockOptimizer(m_optimizerMutex);
    updatedPoseVectorGraph = updatePoseFromGraph();  // 404 (optimizeGraph() is commented out at 403)
  }
  {
    std::lock_guard<std::mutex> lockSend(m_sendMutex);
    m_sendPose = updatedPoseVectorGraph;
    m_sendPoseData = true;
  }
  sendPose();
  sendCones();
}

void Slam::sendPose() {
  if (!onSendPose) return;
  std::lock_guard<std::mutex> lockSend(m_sendMutex);
  onSendPose(m_sendPose);
}

void Slam::sendCones() {
  if (!onSendCones) return;
  Vector3d pose;
  {
    std::lock_guard<std::mutex> lockSend(m_sendMutex);
    pose = m_sendPose;
  }
  std::lock_guard<std::mutex> lockMap(m_mapMutex);
  refreshMapFromMirror();
  onSendCones(m_map, m_currentConeIndex, pose);
}

// slam.cpp:656-679 without the OD4 sends: which cones go out, in which order, with which payload.
// The reference computes index = (cci + i < size) ? cci + i : cci + i - size with no further check;
// an index that is still out of range (map smaller than the packet) is skipped here instead of read.
std::vector<ConePacketEntry> Slam::buildConePacket() {
  Vector3d pose;
  {
    std::lock_guard<std::mutex> lockSend(m_sendMutex);
    pose = m_sendPose;
  }
  std::lock_guard<std::mutex> lockMap(m_mapMutex);
  refreshMapFromMirror();
  std::vector<ConePacketEntry> out;
  const size_t size = m_map.size();
  for (uint32_t i = 0; i < m_conesPerPacket; i++) {
    size_t index = (m_currentConeIndex + i < size) ? (m_currentConeIndex + i) : (m_currentConeIndex + i - size);
    if (index >= size) continue;
    ConePacketEntry e;
    e.objectId = i;
    e.mapIndex = (int)index;
    ConeDirection d = m_map[index].getDirection(pose);
    e.azimuthAngle = d.azimuthAngle;
    e.zenithAngle = d.zenithAngle;
    e.distance = m_map[index].getDistance(pose).distance;
    e.type = m_map[index].getType();
    out.push_back(e);
  }
  return out;
}

std::vector<Vector3d> Slam::drawPoses() {
  std::lock_guard<std::mutex> lockSensor(m_sensorMutex);
  return m_poses;
}

std::vector<Cone> Slam::drawCones() {
  std::lock_guard<std::mutex> lock(m_mapMutex);
  refreshMapFromMirror();
  return m_map;
}

Vector3d Slam::drawCurrentPose() {
  if (m_loopClosingComplete) {
    std::lock_guard<std::mutex> lock(m_sendMutex);
    return m_sendPose;
  } else {
    std::lock_guard<std::mutex> lock(m_sensorMutex);
    return m_odometryData;
  }
}

std::vector<std::vector<int>> Slam::drawGraph() {
  std::lock_guard<std::mutex> lock1(m_mapMutex);
  std::lock_guard<std::mutex> lock2(m_sensorMutex);
  return m_connectivityGraph;
}

int Slam::getPoseEstimate(int id, double out[3]) { return slam_b200_graph_get_vertex(m_ctx, id, out); }

// slam_c.cpp -- flat C wrappers around the C++ Slam mirror so the Python tests can drive it the way
// the reference's frame-gathering thread drives Slam::performSLAM.
#include <chrono>
#include <cstring>
#include <exception>
#include <string>

#include "frame_assembler.hpp"
#include "rec_reader.hpp"
#include "slam.hpp"
#include "wgs84.hpp"

namespace {
thread_local std::string g_err;
}

extern "C" {

const char* slamhost_last_error() { return g_err.c_str(); }

// GNSS priors (wgs84.hpp): the conversions of Slam::nextPose / nextSplitPose / sendPose
void slamhost_wgs84_to_cartesian(const double* ref2, const double* pos2, double* out2) { slamwgs84::toCartesian(ref2, pos2, out2); }
void slamhost_wgs84_from_cartesian(const double* ref2, const double* xy2, double* out2) { slamwgs84::fromCartesian(ref2, xy2, out2); }
double slamhost_heading_from_north(float northHeading) { return slamwgs84::headingFromNorth(northHeading); }

void* slamhost_create(double sameConeThreshold, double coneMappingThreshold, int conesPerPacket, int cudaDevice) {
  try {
    std::map<std::string, std::string> args;
    args["gatheringTimeMs"] = "110";
    args["sameConeThreshold"] = std::to_string(sameConeThreshold);
    args["refLatitude"] = "57.70924648";
    args["refLongitude"] = "11.9462";
    args["timeBetweenKeyframes"] = "0.5";
    args["coneMappingThreshold"] = std::to_string(coneMappingThreshold);
    args["conesPerPacket"] = std::to_string(conesPerPacket);
    args["id"] = "120";
    args["cudaDevice"] = std::to_string(cudaDevice);
    return new Slam(args);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// missing-key behaviour of the reference constructor (std::stoi throws, slam.cpp:739-747)
int slamhost_create_missing_key_throws() {
  try {
    std::map<std::string, std::string> args;
    args["sameConeThreshold"] = "1.2";
    Slam s(args);
    return 0;
  } catch (const std::exception&) {
    return 1;
  }
}

void slamhost_destroy(void* h) { delete static_cast<Slam*>(h); }
// opt-in localiser repair, see Slam::setLocalizerRepair
void slamhost_set_localizer_repair(void* h, int on, int window) { static_cast<Slam*>(h)->setLocalizerRepair(on != 0, window); }

// one frame: what initializeCollection hands to performSLAM (slam.cpp:254) after nextPose stored
// the odometry.  Returns the frame kind (-1 rejected, 0 mapping, 1 loop closed, 2 localiser) or -100.
int slamhost_perform(void* h, const double* cones4xN, int n, const double* pose3, float yawRate, double yawElapsed,
                     int32_t* idx, int32_t* status) {
  try {
    Slam& s = *static_cast<Slam*>(h);
    s.setOdometry(pose3[0], pose3[1], pose3[2]);
    s.setYawRate(yawRate, yawElapsed);
    slamtypes::MatrixXd m(4, n);
    if (n) std::memcpy(m.data(), cones4xN, sizeof(double) * 4 * (size_t)n);
    s.performSLAM(m);
    for (int i = 0; i < n; i++) {
      idx[i] = s.lastIdx()[i];
      status[i] = s.lastStatus()[i];
    }
    return s.lastFrameKind();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -100;
  }
}

// A whole drive with the clock INSIDE (steady_clock around performSLAM, like the reference replay harness times
// the reference's performSLAM): out6 = {mapping s, mapping frames, loop-closing s, frames, localiser s, frames}.
// cones_flat: the frames' 4 x ncols[k] column-major matrices one after another.  Returns frames run or -100.
int slamhost_replay_timed(void* h, int nframes, const double* cones_flat, const int32_t* ncols, const double* poses3,
                          double* out6) {
  try {
    Slam& s = *static_cast<Slam*>(h);
    for (int q = 0; q < 6; q++) out6[q] = 0;
    size_t off = 0;
    for (int k = 0; k < nframes; k++) {
      const int n = ncols[k];
      slamtypes::MatrixXd m(4, n);
      if (n) std::memcpy(m.data(), cones_flat + off, sizeof(double) * 4 * (size_t)n);
      off += 4 * (size_t)n;
      s.setOdometry(poses3[3 * k], poses3[3 * k + 1], poses3[3 * k + 2]);
      s.setYawRate(0.0f, 0.0);
      const auto t0 = std::chrono::steady_clock::now();
      s.performSLAM(m);
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      const int kind = s.lastFrameKind();
      if (kind >= 0 && kind <= 2) { out6[2 * kind] += dt; out6[2 * kind + 1] += 1; }
    }
    return nframes;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -100;
  }
}

// [currentConeIndex, poseId, loopClosing, loopClosingComplete, optimizeCalls, lastIterations, nChi2, mapSize]
void slamhost_state(void* h, int* out8) {
  Slam& s = *static_cast<Slam*>(h);
  out8[0] = (int)s.currentConeIndex(); out8[1] = s.poseId(); out8[2] = s.loopClosing();
  out8[3] = s.loopClosingComplete(); out8[4] = s.optimizeCalls(); out8[5] = s.lastIterations();
  out8[6] = (int)s.chi2Log().size(); out8[7] = (int)s.drawCones().size();
}
void slamhost_chi2_log(void* h, double* out) {
  Slam& s = *static_cast<Slam*>(h);
  for (size_t k = 0; k < s.chi2Log().size(); k++) out[k] = s.chi2Log()[k];
}
void slamhost_draw_cones(void* h, double* x, double* y, int* type, int* id) {
  std::vector<Cone> m = static_cast<Slam*>(h)->drawCones();
  for (size_t j = 0; j < m.size(); j++) { x[j] = m[j].getX(); y[j] = m[j].getY(); type[j] = m[j].getType(); id[j] = m[j].getId(); }
}
int slamhost_draw_poses(void* h, double* out3, int cap) {
  std::vector<slamtypes::Vector3d> p = static_cast<Slam*>(h)->drawPoses();
  for (size_t k = 0; k < p.size() && (int)k < cap; k++) { out3[3 * k] = p[k](0); out3[3 * k + 1] = p[k](1); out3[3 * k + 2] = p[k](2); }
  return (int)p.size();
}
void slamhost_draw_current_pose(void* h, double* out3) {
  slamtypes::Vector3d p = static_cast<Slam*>(h)->drawCurrentPose();
  out3[0] = p(0); out3[1] = p(1); out3[2] = p(2);
}
// flattened drawGraph(): counts[k] = edges of pose k; flat = concatenated cone ids; returns number of poses
int slamhost_draw_graph(void* h, int* counts, int cap_counts, int* flat, int cap_flat) {
  std::vector<std::vector<int>> g = static_cast<Slam*>(h)->drawGraph();
  int q = 0;
  for (size_t k = 0; k < g.size(); k++) {
    if ((int)k < cap_counts) counts[k] = (int)g[k].size();
    for (int c : g[k]) { if (q < cap_flat) flat[q] = c; q++; }
  }
  return (int)g.size();
}
int slamhost_pose_estimate(void* h, int id, double* out3) { return static_cast<Slam*>(h)->getPoseEstimate(id, out3); }
// Cone::getDirection / getDistance (src/cone.cpp:34-53)
void slamhost_cone_bearing(double cx, double cy, const double* pose3, float* az, float* dist) {
  Cone c(cx, cy, 1, 0);
  slamtypes::Vector3d p(pose3[0], pose3[1], pose3[2]);
  *az = c.getDirection(p).azimuthAngle;
  *dist = c.getDistance(p).distance;
}

// Slam::sendCones payload; returns the number of entries
int slamhost_cone_packet(void* h, int cap, int* mapIndex, float* az, float* dist, int* type) {
  std::vector<ConePacketEntry> p = static_cast<Slam*>(h)->buildConePacket();
  for (size_t k = 0; k < p.size() && (int)k < cap; k++) {
    mapIndex[k] = p[k].mapIndex; az[k] = p[k].azimuthAngle; dist[k] = p[k].distance; type[k] = p[k].type;
  }
  return (int)p.size();
}

// ---- frame assembler: host logic only, usable without a GPU ----------------
void* frameasm_create(int gatheringTimeMs, double timeBetweenKeyframes) {
  return new FrameAssembler(gatheringTimeMs, timeBetweenKeyframes);
}
void frameasm_destroy(void* h) { delete static_cast<FrameAssembler*>(h); }
void frameasm_add_direction(void* h, unsigned id, float az, float zen, long long now_us) {
  static_cast<FrameAssembler*>(h)->addDirection(id, az, zen, now_us);
}
void frameasm_add_distance(void* h, unsigned id, float d, long long now_us) {
  static_cast<FrameAssembler*>(h)->addDistance(id, d, now_us);
}
void frameasm_add_type(void* h, unsigned id, unsigned type, long long now_us) {
  static_cast<FrameAssembler*>(h)->addType(id, type, now_us);
}
// returns the number of columns of the emitted frame (written column-major into out4xN, capacity
// cap columns) or -1 when no keyframe was emitted
int frameasm_poll(void* h, long long now_us, double* out4xN, int cap) {
  slamtypes::MatrixXd f;
  if (!static_cast<FrameAssembler*>(h)->poll(now_us, f)) return -1;
  int n = (int)f.cols();
  for (int j = 0; j < n && j < cap; j++)
    for (int r = 0; r < 4; r++) out4xN[4 * (size_t)j + r] = f(r, j);
  return n;
}
// [frameOpen, framesGathered, framesDropped, messagesOutOfRange, capacity]
void frameasm_state(void* h, int* out5) {
  FrameAssembler& a = *static_cast<FrameAssembler*>(h);
  out5[0] = a.frameOpen(); out5[1] = a.framesGathered(); out5[2] = a.framesDroppedByKeyframeGate();
  out5[3] = a.messagesOutOfRange(); out5[4] = a.capacity();
}


// ---- recordings (rec_reader.hpp) ----
// Decodes every envelope of a .rec file: dataType, senderStamp, sample/sent time stamps, objectId (cone
// messages) and up to four payload numbers (direction: az, zen; distance: d; type: type; geolocation: lat,
// lon, altitude, heading; wgs84: lat, lon; heading: north; angular velocity: x, y, z).  Returns the number
// of envelopes (arrays filled up to cap); stats3 = {bytes skipped, truncated tail bytes, envelopes}.
int slamrec_read(const char* path, int cap, int32_t* dataType, uint32_t* sender, int64_t* sample_us, int64_t* sent_us,
                 uint32_t* objectId, double* fields4, int64_t* stats3) {
  slamrec::Reader r;
  if (!r.open(path)) { g_err = "cannot open recording"; return -1; }
  slamrec::Envelope e;
  int n = 0;
  while (r.next(e)) {
    if (n < cap) {
      dataType[n] = e.dataType; sender[n] = e.senderStamp; sample_us[n] = e.sample_us; sent_us[n] = e.sent_us;
      uint32_t id = 0, ty = 0;
      float a = 0, b = 0, c = 0;
      double la = 0, lo = 0;
      double* f = fields4 + 4 * (size_t)n;
      f[0] = f[1] = f[2] = f[3] = 0;
      switch (e.dataType) {
        case slamrec::ID_OBJECT_DIRECTION: slamrec::decodeObjectDirection(e.payload, e.payload_len, id, a, b); f[0] = a; f[1] = b; break;
        case slamrec::ID_OBJECT_DISTANCE: slamrec::decodeObjectDistance(e.payload, e.payload_len, id, a); f[0] = a; break;
        case slamrec::ID_OBJECT_TYPE: slamrec::decodeObjectType(e.payload, e.payload_len, id, ty); f[0] = ty; break;
        case slamrec::ID_GEOLOCATION: slamrec::decodeGeolocation(e.payload, e.payload_len, la, lo, a, b); f[0] = la; f[1] = lo; f[2] = a; f[3] = b; break;
        case slamrec::ID_GEODETIC_WGS84: slamrec::decodeGeodeticWgs84(e.payload, e.payload_len, la, lo); f[0] = la; f[1] = lo; break;
        case slamrec::ID_GEODETIC_HEADING: slamrec::decodeGeodeticHeading(e.payload, e.payload_len, a); f[0] = a; break;
        case slamrec::ID_ANGULAR_VELOCITY: slamrec::decodeAngularVelocity(e.payload, e.payload_len, a, b, c); f[0] = a; f[1] = b; f[2] = c; break;
        default: break;
      }
      objectId[n] = id;
    }
    n++;
  }
  if (stats3) { stats3[0] = (int64_t)r.bytesSkipped(); stats3[1] = (int64_t)r.truncatedTail(); stats3[2] = (int64_t)r.envelopesRead(); }
  return n;
}

// Replays a recording through the front half (slamrec::replay).  Frames are returned flattened: ncols[k]
// columns of frame k at cones + 4 * (sum of earlier ncols); odo3 / yaw / yawElapsed / time_us per frame.
// stats9 = envelopes, coneMessages, poseMessages, ignoredSender, ignoredType, malformed, framesGathered,
// framesDroppedByKeyframeGate, framesEmitted.  Returns the number of frames emitted.
int slamrec_replay(const char* path, uint32_t detectConeId, uint32_t estimationId, int gatheringTimeMs,
                   double timeBetweenKeyframes, double refLat, double refLon, int capFrames, int capCols, int32_t* ncols,
                   double* cones, double* odo3, float* yaw, double* yawElapsed, int64_t* time_us, int64_t* stats9) {
  slamrec::Reader r;
  if (!r.open(path)) { g_err = "cannot open recording"; return -1; }
  slamrec::ReplayConfig cfg;
  cfg.detectConeId = detectConeId; cfg.estimationId = estimationId; cfg.gatheringTimeMs = gatheringTimeMs;
  cfg.timeBetweenKeyframes = timeBetweenKeyframes; cfg.refLatitude = refLat; cfg.refLongitude = refLon;
  int nf = 0;
  long used = 0;
  slamrec::ReplayStats st = slamrec::replay(r, cfg, [&](const slamrec::ReplayFrame& f) {
    const int n = (int)f.cones.cols();
    if (nf < capFrames && used + n <= capCols) {
      ncols[nf] = n;
      for (int j = 0; j < n; j++)
        for (int q = 0; q < 4; q++) cones[4 * (used + j) + q] = f.cones(q, j);
      odo3[3 * nf] = f.odometry[0]; odo3[3 * nf + 1] = f.odometry[1]; odo3[3 * nf + 2] = f.odometry[2];
      yaw[nf] = f.yawRate; yawElapsed[nf] = f.yawElapsed; time_us[nf] = f.time_us;
      used += n;
    }
    nf++;
  });
  if (stats9) {
    const int64_t v[9] = {(int64_t)st.envelopes, (int64_t)st.coneMessages, (int64_t)st.poseMessages, (int64_t)st.ignoredSender,
                          (int64_t)st.ignoredType, (int64_t)st.malformed, st.framesGathered, st.framesDroppedByKeyframeGate, st.framesEmitted};
    for (int k = 0; k < 9; k++) stats9[k] = v[k];
  }
  return nf;
}

// Recording -> front half on recorded time -> the drop-in Slam: every keyframe the replay releases goes
// through setOdometry / setYawRate / performSLAM exactly as Slam::initializeCollection would hand it over
// (slam.cpp:248-255).  Returns the number of frames performed, -1 if the file cannot be read, -100 on error.
int slamrec_replay_into_slam(const char* path, void* slam, uint32_t detectConeId, uint32_t estimationId, int gatheringTimeMs,
                             double timeBetweenKeyframes, double refLat, double refLon) {
  try {
    slamrec::Reader r;
    if (!r.open(path)) { g_err = "cannot open recording"; return -1; }
    slamrec::ReplayConfig cfg;
    cfg.detectConeId = detectConeId; cfg.estimationId = estimationId; cfg.gatheringTimeMs = gatheringTimeMs;
    cfg.timeBetweenKeyframes = timeBetweenKeyframes; cfg.refLatitude = refLat; cfg.refLongitude = refLon;
    Slam& s = *static_cast<Slam*>(slam);
    int n = 0;
    slamrec::replay(r, cfg, [&](const slamrec::ReplayFrame& f) {
      s.setOdometry(f.odometry[0], f.odometry[1], f.odometry[2]);
      s.setYawRate(f.yawRate, f.yawElapsed);
      s.performSLAM(f.cones);
      n++;
    });
    return n;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -100;
  }
}

}  // extern "C"

// integration/public_api_harness.cpp -- TEST INFRASTRUCTURE.  Drives `class Slam` of the reference tree through
// its PUBLIC interface only (slam.hpp:52-62): the constructor taking the command-line map and a cluon::OD4Session,
// nextPose / nextYawRate / nextCone fed with cluon::data::Envelope objects exactly as the data triggers of
// opendlv-logic-cfsd18-sensation-slam.cpp:71-108 deliver them, and drawCones / drawPoses / drawCurrentPose /
// drawGraph to read the result.  The same source is compiled twice (integration/build.sh):
//   * against the reference's UNMODIFIED src/slam.hpp + slam.cpp (g2o answered by oracle/g2o_facade) ->
//     integration/_build/ref_public_replay: the reference behaviour, run on the CPU to make the golden file;
//   * against the PATCHED copy (integration/slam_b200.patch applied) linked with libslam_b200.so ->
//     integration/_build/patched_public_replay: the drop-in, run on the B200 and compared with the golden file.
// Nothing private is touched: no `#define private public`, no friend.
//
// The reference's front half runs on the wall clock (a detached thread per frame busy-waits gatheringTimeMs, then
// isKeyframe() compares the time since the last keyframe in MILLISECONDS with the raw timeBetweenKeyframes setting),
// so the harness paces the frames: all envelopes of a frame go in at once, then it waits until the frame's pose has
// appeared in drawPoses() and drawCones() has got past the map mutex the frame holds while it is being processed.
//
// input  (binary, argv[1]): the extended record format of oracle/ref_slam_replay.cpp: int32 -nframes, double
//         sameConeThreshold, double coneMappingThreshold, then per frame float yawRate, int64 elapsed_us, double pose[3],
//         int32 N, double cones[4 N].  A column whose four entries are all zero is NOT sent (an absent objectId).
// output (text, argv[2]): per frame "F k <poses> <graph rows>"; at the end "M j x y type id" per map cone, "P k x y h" per
//         stored pose, "G k n id..." per row of drawGraph(), "C x y h" = drawCurrentPose(), "D n" frames the class dropped, "S n" frames whose messages took more than half a
//         gathering window to send (the run is then repeated); doubles as hex floats.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "cluon-complete.hpp"
#include "opendlv-standard-message-set.hpp"
#include "slam.hpp"
#include "WGS84toCartesian.hpp"

namespace {
template <class M>
cluon::data::Envelope envelope(M& msg, const cluon::data::TimeStamp& ts, uint32_t sender) {
  cluon::ToProtoVisitor v;
  msg.accept(v);
  cluon::data::Envelope e;
  e.dataType(M::ID()).serializedData(v.encodedData()).sent(ts).received(ts).sampleTimeStamp(ts).senderStamp(sender);
  return e;
}
cluon::data::TimeStamp stamp(int64_t us) {
  cluon::data::TimeStamp t;
  t.seconds(static_cast<int32_t>(us / 1000000)).microseconds(static_cast<int32_t>(us % 1000000));
  return t;
}
void sleep_us(int us) { std::this_thread::sleep_for(std::chrono::microseconds(us)); }
// (latitude, longitude) whose wgs84::toCartesian -- what Slam::nextPose applies -- is (x, y): the header's own
// fromCartesian is a coarse stepping inverse (metres off), so it only seeds a Newton iteration on toCartesian
std::array<double, 2> latlon_for(const std::array<double, 2>& ref, double x, double y) {
  std::array<double, 2> ll = wgs84::fromCartesian(ref, std::array<double, 2>{x, y});
  for (int it = 0; it < 8; it++) {
    const std::array<double, 2> f = wgs84::toCartesian(ref, ll);
    const double h = 1e-7;
    const std::array<double, 2> fa = wgs84::toCartesian(ref, std::array<double, 2>{ll[0] + h, ll[1]});
    const std::array<double, 2> fb = wgs84::toCartesian(ref, std::array<double, 2>{ll[0], ll[1] + h});
    const double a = (fa[0] - f[0]) / h, b = (fb[0] - f[0]) / h, c = (fa[1] - f[1]) / h, d = (fb[1] - f[1]) / h;
    const double det = a * d - b * c, rx = x - f[0], ry = y - f[1];
    if (det == 0) break;
    ll[0] += (d * rx - b * ry) / det;
    ll[1] += (-c * rx + a * ry) / det;
    if (std::fabs(rx) + std::fabs(ry) < 1e-10) break;
  }
  return ll;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s drive.bin out.txt [gatheringTimeMs]\n", argv[0]); return 2; }
  std::ifstream in(argv[1], std::ios::binary);
  if (!in) return 2;
  const int gathering_ms = argc > 3 ? std::atoi(argv[3]) : 10;
  int32_t nframes = 0;
  double thr = 0, map_thr = 0;
  in.read(reinterpret_cast<char*>(&nframes), 4);
  in.read(reinterpret_cast<char*>(&thr), 8);
  in.read(reinterpret_cast<char*>(&map_thr), 8);
  if (nframes >= 0) { std::fprintf(stderr, "extended records expected\n"); return 2; }
  nframes = -nframes;
  const std::array<double, 2> ref = {57.70924648, 11.9462};
  std::map<std::string, std::string> args;
  args["gatheringTimeMs"] = std::to_string(gathering_ms);
  args["sameConeThreshold"] = std::to_string(thr);
  args["refLatitude"] = "57.70924648";
  args["refLongitude"] = "11.9462";
  args["timeBetweenKeyframes"] = "0.5";
  args["coneMappingThreshold"] = std::to_string(map_thr);
  args["conesPerPacket"] = "20";
  args["id"] = "120";
  FILE* out = std::fopen(argv[2], "w");
  if (!out) return 2;
  // the reference narrates every observation on std::cout
  std::ofstream devnull("/dev/null");
  std::streambuf* saved = std::cout.rdbuf(devnull.rdbuf());
  int dropped = 0, suspect = 0;
  {
    cluon::OD4Session od4{111};
    Slam slam(args, od4);
    int64_t t_us = 1000LL * 1000000LL;  // sample time of the first frame
    for (int k = 0; k < nframes; k++, t_us += 100000) {
      float yaw = 0;
      int64_t elapsed = 0;
      double pose[3];
      int32_t n = 0;
      in.read(reinterpret_cast<char*>(&yaw), 4);
      in.read(reinterpret_cast<char*>(&elapsed), 8);
      in.read(reinterpret_cast<char*>(pose), 24);
      in.read(reinterpret_cast<char*>(&n), 4);
      std::vector<double> cones(4 * static_cast<size_t>(n));
      in.read(reinterpret_cast<char*>(cones.data()), static_cast<std::streamsize>(cones.size() * 8));
      if (!in) { std::fprintf(stderr, "short read at frame %d\n", k); return 2; }
      const size_t before = slam.drawPoses().size();
      {  // opendlv.proxy.AngularVelocityReading -> nextYawRate (divides by 4)
        opendlv::proxy::AngularVelocityReading m;
        m.angularVelocityZ(yaw * 4.0f);
        slam.nextYawRate(envelope(m, stamp(t_us - elapsed), 112));
      }
      {  // opendlv.logic.sensation.Geolocation -> nextPose (WGS84 -> Cartesian about the reference point)
        const std::array<double, 2> ll = latlon_for(ref, pose[0], pose[1]);
        opendlv::logic::sensation::Geolocation m;
        m.latitude(ll[0]).longitude(ll[1]).heading(static_cast<float>(pose[2]));
        slam.nextPose(envelope(m, stamp(t_us), 112));
      }
      int sent = 0;
      const auto send0 = std::chrono::steady_clock::now();
      for (int32_t i = 0; i < n; i++) {
        const double* c = &cones[4 * static_cast<size_t>(i)];
        if (c[0] == 0 && c[1] == 0 && c[2] == 0 && c[3] == 0) continue;  // absent objectId: the column stays zero
        opendlv::logic::perception::ObjectDirection d;
        d.objectId(static_cast<uint32_t>(i)).azimuthAngle(static_cast<float>(c[0])).zenithAngle(static_cast<float>(c[1]));
        slam.nextCone(envelope(d, stamp(t_us), 116));
        opendlv::logic::perception::ObjectDistance r;
        r.objectId(static_cast<uint32_t>(i)).distance(static_cast<float>(c[2]));
        slam.nextCone(envelope(r, stamp(t_us), 116));
        opendlv::logic::perception::ObjectType ty;
        ty.objectId(static_cast<uint32_t>(i)).type(static_cast<uint32_t>(c[3]));
        slam.nextCone(envelope(ty, stamp(t_us), 116));
        sent++;
      }
      // All messages of a frame must land inside ONE gathering window of the reference's wall-clock collector: a harness
      // thread that lost the CPU for half a window in the middle of a frame may have split it -- the run is marked
      // suspect and the caller repeats it.
      if (std::chrono::steady_clock::now() - send0 > std::chrono::microseconds(500 * gathering_ms)) suspect++;
      // wait for the frame: its pose appears (performSLAM, slam.cpp:320), then the map mutex is free again
      bool seen = false;
      if (sent > 0) {
        const auto t0 = std::chrono::steady_clock::now();
        while (std::chrono::steady_clock::now() - t0 < std::chrono::milliseconds(gathering_ms + 150)) {
          if (slam.drawPoses().size() > before) { seen = true; break; }
          sleep_us(100);
        }
      }
      if (seen) {
        // drawCones() takes the map mutex addConesToMap holds for the whole frame (the loop-closing frame: through its
        // optimise burst).  (Polling drawGraph() instead is not an option: the reference grows m_connectivityGraph
        // under the optimizer mutex only while drawGraph() copies it under map + sensor mutex -- the harness
        // crashed the unmodified reference that way, free(): invalid pointer.)
        for (int rep = 0; rep < 3; rep++) {
          sleep_us(700);
          (void)slam.drawCones();
        }
            } else {
        dropped++;  // behind the 200 m gate of performSLAM (slam.cpp:300-303) or an empty frame
        sleep_us(1000 * gathering_ms + 2000);
      }
      std::fprintf(out, "F %d %zu %zu\n", k, slam.drawPoses().size(), slam.drawGraph().size());
    }
    sleep_us(20000);
    const std::vector<Cone> map = slam.drawCones();
    for (size_t j = 0; j < map.size(); j++) {
      Cone c = map[j];
      std::fprintf(out, "M %zu %a %a %d %d\n", j, c.getX(), c.getY(), c.getType(), c.getId());
    }
    const std::vector<Eigen::Vector3d> poses = slam.drawPoses();
    for (size_t k = 0; k < poses.size(); k++) std::fprintf(out, "P %zu %a %a %a\n", k, poses[k](0), poses[k](1), poses[k](2));
    const std::vector<std::vector<int>> graph = slam.drawGraph();
    for (size_t k = 0; k < graph.size(); k++) {
      std::fprintf(out, "G %zu %zu", k, graph[k].size());
      for (int id : graph[k]) std::fprintf(out, " %d", id);
      std::fprintf(out, "\n");
    }
    const Eigen::Vector3d cur = slam.drawCurrentPose();
    std::fprintf(out, "C %a %a %a\n", cur(0), cur(1), cur(2));
    std::fprintf(out, "D %d\n", dropped);
    std::fprintf(out, "S %d\n", suspect);
  }
  std::fclose(out);
  std::cout.rdbuf(saved);
  return 0;
}

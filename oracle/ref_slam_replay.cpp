// oracle/ref_slam_replay.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Drives the reference's REAL `class Slam` (compiled from /root/reference/src/slam.cpp + cone.cpp where
// they lie, g2o replaced by oracle/g2o_facade) through its back half, frame by frame, and records what
// it decides: which map cones every frame was associated with (the row performSLAM appends to
// m_connectivityGraph), map size, current-cone index, the loop-closure flags, the pose it stored, and at
// the end the map, the stored poses and the optimised pose vertices.
//
// Slam::performSLAM is private and normally reached from a detached, wall-clock-gated thread
// (slam.cpp:94,221-257); the harness reaches it with the usual test trick of re-declaring `private`
// while including slam.hpp -- the reference sources are not modified.  m_odometryData is what
// Slam::nextPose would have stored (slam.cpp:207-209); the yaw/geolocation stamps stay at their defaults,
// so the heading correction of performSLAM (309-318) is inactive, as in the oracle's replay.
//
// input  (binary, argv[1]): int32 nframes, double sameConeThreshold, double coneMappingThreshold, then per
//         frame: double pose[3], int32 N, double cones[4*N] (column-major az, zen, range, type).
//         nframes < 0 = extended records: every frame additionally carries float yawRate, int64 elapsed_us
//         BEFORE the pose; they become m_yawRate and the distance between m_yawReceivedTime and m_lastTimeStamp,
//         which drives the heading correction of performSLAM (slam.cpp:309-318).
// output (text, argv[2]):   one line per frame + a trailer, parsed by tests/golden/make_c1_reference_replay.py
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include "g2o_facade.hpp"
#include <Eigen/Dense>
#include "cluon-complete.hpp"
#include "opendlv-standard-message-set.hpp"
#include "cone.hpp"

#define private public
#include "slam.hpp"
#undef private

// `ref_slam_replay --conv in.txt out.txt`: the reference's private conversion helpers on their own.  Input
// lines "az zen range type px py ptheta" (hex floats); output per line: transformConeToCoG (2), Spherical2Cartesian
// (3), coneToGlobal (3), all as hex floats.
static int conversions(const char* fin, const char* fout) {
  std::map<std::string, std::string> args;
  args["gatheringTimeMs"] = "110"; args["sameConeThreshold"] = "1.2"; args["refLatitude"] = "57.70924648";
  args["refLongitude"] = "11.9462"; args["timeBetweenKeyframes"] = "0.5"; args["coneMappingThreshold"] = "50";
  args["conesPerPacket"] = "20"; args["id"] = "120";
  std::ofstream devnull("/dev/null");
  std::streambuf* saved = std::cout.rdbuf(devnull.rdbuf());
  cluon::OD4Session od4{111};
  FILE* in = std::fopen(fin, "r");
  FILE* out = std::fopen(fout, "w");
  if (!in || !out) return 2;
  {
    Slam slam(args, od4);
    double az, zen, rng, type, px, py, pt;
    while (std::fscanf(in, " %la %la %la %la %la %la %la", &az, &zen, &rng, &type, &px, &py, &pt) == 7) {
      const Eigen::Vector2d cog = slam.transformConeToCoG(az, rng);
      const Eigen::Vector3d xyz = slam.Spherical2Cartesian(az, zen, rng);
      Eigen::MatrixXd col(4, 1);
      col << az, zen, rng, type;
      const Eigen::Vector3d g = slam.coneToGlobal(Eigen::Vector3d(px, py, pt), col);
      std::fprintf(out, "%a %a %a %a %a %a %a %a\n", cog(0), cog(1), xyz(0), xyz(1), xyz(2), g(0), g(1), g(2));
    }
  }
  std::fclose(in);
  std::fclose(out);
  std::cout.rdbuf(saved);
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 4 && std::string(argv[1]) == "--conv") return conversions(argv[2], argv[3]);
  if (argc < 3) { std::fprintf(stderr, "usage: ref_slam_replay frames.bin out.txt | --conv in.txt out.txt\n"); return 2; }
  std::ifstream in(argv[1], std::ios::binary);
  int32_t nframes = 0;
  double thr = 0, mapThr = 0;
  in.read(reinterpret_cast<char*>(&nframes), 4);
  const bool extended = nframes < 0;
  if (extended) nframes = -nframes;
  in.read(reinterpret_cast<char*>(&thr), 8);
  in.read(reinterpret_cast<char*>(&mapThr), 8);
  std::map<std::string, std::string> args;
  args["cid"] = "111";
  args["gatheringTimeMs"] = "110";
  args["sameConeThreshold"] = std::to_string(thr);
  args["refLatitude"] = "57.70924648";
  args["refLongitude"] = "11.9462";
  args["timeBetweenKeyframes"] = "0.5";
  args["coneMappingThreshold"] = std::to_string(mapThr);
  args["conesPerPacket"] = "20";
  args["id"] = "120";
  // the reference prints per observation (slam.cpp:590 and friends): silence stdout for the replay
  std::ofstream devnull("/dev/null");
  std::streambuf* saved = std::cout.rdbuf(devnull.rdbuf());
  cluon::OD4Session od4{111};
  FILE* out = std::fopen(argv[2], "w");
  {
    Slam slam(args, od4);
    double t_mapping = 0, t_closing = 0, t_localise = 0;  // seconds inside performSLAM, by frame kind
    int n_mapping = 0, n_closing = 0, n_localise = 0;
    for (int f = 0; f < nframes; f++) {
      double pose[3];
      int32_t N = 0;
      if (extended) {
        float yawRate = 0;
        int64_t elapsed_us = 0;
        in.read(reinterpret_cast<char*>(&yawRate), 4);
        in.read(reinterpret_cast<char*>(&elapsed_us), 8);
        slam.m_yawRate = yawRate;                                   // what Slam::nextYawRate stores (slam.cpp:216)
        cluon::data::TimeStamp zero, later;
        later.seconds((int32_t)(elapsed_us / 1000000)).microseconds((int32_t)(elapsed_us % 1000000));
        slam.m_yawReceivedTime = zero;                              // 217
        slam.m_lastTimeStamp = later;                               // nextCone, 73 / 102 / 129
      }
      in.read(reinterpret_cast<char*>(pose), 24);
      in.read(reinterpret_cast<char*>(&N), 4);
      Eigen::MatrixXd cones(4, N);
      if (N) in.read(reinterpret_cast<char*>(cones.data()), (std::streamsize)sizeof(double) * 4 * N);
      slam.m_odometryData << pose[0], pose[1], pose[2];
      const size_t rows0 = slam.m_connectivityGraph.size();
      const bool closedBefore = slam.m_loopClosingComplete;
      const auto c0 = std::chrono::steady_clock::now();
      slam.performSLAM(cones);
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count();
      if (closedBefore) { t_localise += dt; n_localise++; }
      else if (slam.m_loopClosingComplete) { t_closing += dt; n_closing++; }
      else { t_mapping += dt; n_mapping++; }
      std::fprintf(out, "F %d %zu %u %d %d %d", f, slam.m_map.size(), slam.m_currentConeIndex, (int)slam.m_loopClosing,
                   (int)slam.m_loopClosingComplete, slam.m_poseId);
      if (slam.m_connectivityGraph.size() > rows0) {
        const std::vector<int>& row = slam.m_connectivityGraph.back();
        std::fprintf(out, " %zu", row.size());
        for (int id : row) std::fprintf(out, " %d", id);
      } else {
        std::fprintf(out, " -1");
      }
      std::fprintf(out, "\n");
    }
    for (size_t j = 0; j < slam.m_map.size(); j++)
      std::fprintf(out, "M %zu %a %a %d %d\n", j, slam.m_map[j].getX(), slam.m_map[j].getY(), slam.m_map[j].getType(),
                   slam.m_map[j].getId());
    for (size_t k = 0; k < slam.m_poses.size(); k++)
      std::fprintf(out, "P %zu %a %a %a\n", k, slam.m_poses[k](0), slam.m_poses[k](1), slam.m_poses[k](2));
    for (int id = 1000; id < slam.m_poseId; id++) {
      g2o::VertexSE2* v = static_cast<g2o::VertexSE2*>(slam.m_optimizer.vertex(id));
      const Eigen::Vector3d e = v->estimate().toVector();
      std::fprintf(out, "V %d %a %a %a\n", id, e(0), e(1), e(2));
    }
    for (double c : slam.m_optimizer.last_chi2_) std::fprintf(out, "C %a\n", c);
    std::fprintf(out, "S %a %a %a\n", slam.m_sendPose(0), slam.m_sendPose(1), slam.m_sendPose(2));
    // wall time inside performSLAM (stdout of the reference's per-observation prints goes to /dev/null)
    std::fprintf(out, "T %d %.9f %d %.9f %d %.9f\n", n_mapping, t_mapping, n_closing, t_closing, n_localise, t_localise);
  }
  std::fclose(out);
  std::cout.rdbuf(saved);
  return 0;
}

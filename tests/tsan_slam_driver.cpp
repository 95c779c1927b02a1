// tsan_slam_driver.cpp -- test infrastructure: the drop-in Slam class (csrc/host/slam.cpp) driven like the
// reference drives it -- one thread feeding frames through performSLAM (mapping, loop closure, localiser, with
// and without the opt-in repair), a viewer thread hammering drawCones / drawPoses / drawCurrentPose /
// drawGraph / buildConePacket (viewerbuild/src/drawer.cpp does exactly that) -- over the stub backend, under
// ThreadSanitizer.  Exit 0 and "ok" = no data race, no lock-order inversion, no deadlock.
#include <atomic>
#include <cstdio>
#include <map>
#include <string>
#include <thread>

#include "slam.hpp"

int main() {
  for (int repair = 0; repair < 2; repair++) {
    std::map<std::string, std::string> args = {{"gatheringTimeMs", "10"}, {"sameConeThreshold", "1.2"}, {"refLatitude", "57.7"},
                                               {"refLongitude", "11.9"}, {"timeBetweenKeyframes", "0.5"},
                                               {"coneMappingThreshold", "50"}, {"conesPerPacket", "20"}, {"id", "120"}};
    if (repair) { args["localizerRepair"] = "1"; args["localizerWindow"] = "4"; }
    Slam slam(args);
    std::atomic<int> sends{0};
    slam.onSendPose = [&](const slamtypes::Vector3d&) { sends++; };
    slam.onSendCones = [&](const std::vector<Cone>& m, uint32_t, const slamtypes::Vector3d&) { sends += (int)m.size() > 0; };
    std::atomic<bool> done{false};
    long seen = 0;
    std::atomic<int> looks{0};
    std::thread viewer([&] {
      while (!done.load()) {
        looks++;
        seen += (long)slam.drawCones().size() + (long)slam.drawPoses().size() + (long)slam.drawGraph().size();
        seen += (long)slam.buildConePacket().size();
        seen += slam.drawCurrentPose()(0) > 1e300;
      }
    });
    for (int k = 0; k < 400 || (looks.load() < 200 && k < 200000); k++) {   // until the viewer has really looked
      slamtypes::MatrixXd cones(4, 3 + k % 5);
      for (int j = 0; j < cones.cols(); j++) { cones(0, j) = 5 + j; cones(1, j) = 0; cones(2, j) = 4 + j; cones(3, j) = 1 + (j & 1); }
      slam.setOdometry(0.1 * k, 0.05 * k, 0.01 * k);
      slam.setYawRate(0.01f, 0.02);
      slam.performSLAM(cones);
    }
    done = true;
    viewer.join();
    if (!slam.loopClosingComplete() || sends.load() == 0 || seen == 0) { printf("driver did not reach the localiser\n"); return 1; }
  }
  printf("ok\n");
  return 0;
}

// frame_assembler.hpp -- deterministic, clock-injectable restatement of the reference's frame
// gathering and keyframing: Slam::nextCone (src/slam.cpp:67-152),
// Slam::initializeCollection (221-257) and Slam::isKeyframe (286-295).
//
// The reference writes ObjectDirection / ObjectDistance / ObjectType fields into a 4 x capacity
// collector indexed by objectId, and the FIRST message of a frame spawns a detached thread that
// busy-waits gatheringTimeMs of wall clock before snapshotting columns 0..lastObjectId.  Here the
// same state machine is driven by explicit timestamps (microseconds): add*() mirror nextCone(),
// poll(now) closes the frame once the gathering window has elapsed and applies the keyframe gate.
// No threads, no wall clock: replays are reproducible and the frame lands directly in the buffer
// the association kernel takes (column-major 4 x N doubles).
//
// Deliberate deviation: the reference indexes the collector without a bounds check (objectId >=
// capacity is undefined behaviour, slam.cpp:83-84,108,136); here such a message is dropped and counted.
#pragma once
#include <cstdint>
#include <vector>

#include "slam_types.hpp"

class FrameAssembler {
 public:
  // gatheringTimeMs: slam.hpp:99 (m_timeDiffMilliseconds); timeBetweenKeyframes: slam.hpp:115
  FrameAssembler(int32_t gatheringTimeMs, double timeBetweenKeyframes);

  // Slam::nextCone branches (float wire fields widened to double, slam.cpp:83-84,108,136)
  void addDirection(uint32_t objectId, float azimuthAngle, float zenithAngle, int64_t now_us);
  void addDistance(uint32_t objectId, float distance, int64_t now_us);
  void addType(uint32_t objectId, uint32_t type, int64_t now_us);

  // Slam::initializeCollection + isKeyframe: returns true and fills `frame` (4 x N) when the frame
  // that was open has been gathered for longer than gatheringTimeMs AND passes the keyframe gate.
  // A gathered frame that fails the gate is discarded (the reference drops it too, 251-255).
  bool poll(int64_t now_us, slamtypes::MatrixXd& frame);

  bool frameOpen() const { return !m_newFrame; }
  int framesGathered() const { return m_framesGathered; }
  int framesDroppedByKeyframeGate() const { return m_framesDropped; }
  int messagesOutOfRange() const { return m_outOfRange; }
  int capacity() const { return m_capacity; }

 private:
  void touch(uint32_t objectId, int64_t now_us);
  bool isKeyframe(int64_t now_us);

  int32_t m_timeDiffMilliseconds;
  double m_timeBetweenKeyframes;
  std::vector<double> m_coneCollector;  // column-major 4 x m_capacity
  int m_capacity = 100;                 // 4 x 100 at construction (slam.cpp:46), 4 x 1000 after every frame (244)
  uint32_t m_lastObjectId = 0;
  bool m_newFrame = true;
  int64_t m_frameStart_us = 0;
  int64_t m_keyframeTimeStamp_us = 0;   // cluon::data::TimeStamp{} == 0 (slam.cpp:39)
  int m_framesGathered = 0, m_framesDropped = 0, m_outOfRange = 0;
};

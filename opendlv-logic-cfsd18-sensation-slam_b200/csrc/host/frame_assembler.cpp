// frame_assembler.cpp -- see frame_assembler.hpp.
#include "frame_assembler.hpp"

#include <cmath>

FrameAssembler::FrameAssembler(int32_t gatheringTimeMs, double timeBetweenKeyframes)
    : m_timeDiffMilliseconds(gatheringTimeMs), m_timeBetweenKeyframes(timeBetweenKeyframes),
      m_coneCollector(4 * 100, 0.0) {}

// common tail of the three nextCone branches: track the highest object id and open the frame on its
// first message (slam.cpp:81,87-97)
void FrameAssembler::touch(uint32_t objectId, int64_t now_us) {
  m_lastObjectId = (m_lastObjectId < objectId) ? objectId : m_lastObjectId;
  if (m_newFrame) {
    m_newFrame = false;
    m_frameStart_us = now_us;  // the reference starts its busy-wait here (slam.cpp:94, 225)
  }
}

void FrameAssembler::addDirection(uint32_t objectId, float azimuthAngle, float zenithAngle, int64_t now_us) {
  if ((int64_t)objectId >= m_capacity) { m_outOfRange++; return; }
  m_coneCollector[4 * (size_t)objectId + 0] = azimuthAngle;  // slam.cpp:83
  m_coneCollector[4 * (size_t)objectId + 1] = zenithAngle;   // slam.cpp:84
  touch(objectId, now_us);
}

void FrameAssembler::addDistance(uint32_t objectId, float distance, int64_t now_us) {
  if ((int64_t)objectId >= m_capacity) { m_outOfRange++; return; }
  m_coneCollector[4 * (size_t)objectId + 2] = distance;      // slam.cpp:108
  touch(objectId, now_us);
}

void FrameAssembler::addType(uint32_t objectId, uint32_t type, int64_t now_us) {
  if ((int64_t)objectId >= m_capacity) { m_outOfRange++; return; }
  m_coneCollector[4 * (size_t)objectId + 3] = type;          // slam.cpp:136
  touch(objectId, now_us);
}

// slam.cpp:286-295.  timeElapsed is |delta| in MILLIseconds (microseconds / 1000) compared against
// timeBetweenKeyframes as is -- the reference's unit mix-up is reproduced, not repaired.
bool FrameAssembler::isKeyframe(int64_t now_us) {
  double timeElapsed = std::fabs(static_cast<double>(m_keyframeTimeStamp_us - now_us)) / 1000;
  if (timeElapsed > m_timeBetweenKeyframes) {
    m_keyframeTimeStamp_us = now_us;
    return true;
  }
  return false;
}

bool FrameAssembler::poll(int64_t now_us, slamtypes::MatrixXd& frame) {
  if (m_newFrame) return false;                                        // no frame open
  if (!(now_us - m_frameStart_us > (int64_t)m_timeDiffMilliseconds * 1000)) return false;  // 231
  const int n = (int)m_lastObjectId + 1;                               // leftCols(m_lastObjectId+1), 241
  slamtypes::MatrixXd extracted(4, n);
  for (int j = 0; j < n; j++)
    for (int r = 0; r < 4; r++) extracted(r, j) = m_coneCollector[4 * (size_t)j + r];
  m_newFrame = true;                                                   // 242
  m_lastObjectId = 0;                                                  // 243
  m_capacity = 1000;                                                   // 244
  m_coneCollector.assign(4 * 1000, 0.0);
  m_framesGathered++;
  if (n > 0 && isKeyframe(now_us)) {                                   // 248-255
    frame = extracted;
    return true;
  }
  m_framesDropped++;
  return false;
}

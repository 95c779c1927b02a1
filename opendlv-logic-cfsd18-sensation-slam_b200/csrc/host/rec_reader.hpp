// rec_reader.hpp -- reader for cluon `.rec` recordings / OD4 envelope streams and a deterministic replay of
// the reference's front half: the on-disk / on-wire format next to the hot path.
//
// Format (reference: src/cluon-complete-build.hpp, serializeEnvelope 6868-6898 / extractEnvelope 6900-6960):
//   0x0D 0xA4 LEN0 LEN1 LEN2 | Proto-encoded cluon::data::Envelope          (LEN little-endian, 24 bit)
// Envelope (3910-3976): 1 dataType sint32 | 2 serializedData bytes | 3 sent | 4 received | 5 sampleTimeStamp
// (TimeStamp 3746-3784: 1 seconds sint32, 2 microseconds sint32) | 6 senderStamp uint32.  Encoding is the
// protobuf wire format as cluon's ToProtoVisitor writes it (9842-9990): key = id << 3 | wire type, signed
// integers zig-zag varints, float / double fixed32 / fixed64 little-endian, nested messages length-delimited.
// Payloads decoded here are the seven messages Slam consumes (src/opendlv-logic-cfsd18-sensation-slam.cpp:
// 100-106 and opendlv-standard-message-set-v0.9.5.odvd): ObjectDirection 1133, ObjectDistance 1134,
// ObjectType 1131, Geolocation 1116, GeodeticWgs84Reading 19, GeodeticHeadingReading 1051,
// AngularVelocityReading 1031.  Pinned by cluon's own serializer: tests/golden/make_rec_golden.py writes the
// fixture with the reference's cluon + generated message set; tests/test_rec_reader.py reads it back.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "frame_assembler.hpp"
#include "slam_types.hpp"

namespace slamrec {

enum MessageId : int32_t {
  ID_GEODETIC_WGS84 = 19, ID_ANGULAR_VELOCITY = 1031, ID_GEODETIC_HEADING = 1051, ID_GEOLOCATION = 1116,
  ID_OBJECT_TYPE = 1131, ID_OBJECT_DIRECTION = 1133, ID_OBJECT_DISTANCE = 1134
};

struct Envelope {
  int32_t dataType = 0;
  uint32_t senderStamp = 0;
  int64_t sent_us = 0, received_us = 0, sample_us = 0;  // seconds * 1e6 + microseconds (cluon::time::toMicroseconds)
  const uint8_t* payload = nullptr;                       // into the reader's buffer, valid until the next call
  size_t payload_len = 0;
};

// Sequential reader over a byte buffer (a whole .rec file or a captured UDP stream).  next() returns false at
// the end; bytes that do not start an envelope are skipped one at a time like cluon's Player does on a
// corrupt file (resync on the 0x0D 0xA4 header), and counted.
class Reader {
 public:
  Reader() {}
  bool open(const std::string& path);                     // reads the whole file
  void attach(const uint8_t* data, size_t len);            // or parse caller-owned memory
  bool next(Envelope& e);
  size_t envelopesRead() const { return m_count; }
  size_t bytesSkipped() const { return m_skipped; }
  size_t truncatedTail() const { return m_truncated; }    // bytes of an incomplete envelope at the end

 private:
  std::vector<uint8_t> m_file;
  const uint8_t* m_p = nullptr;
  size_t m_len = 0, m_off = 0, m_count = 0, m_skipped = 0, m_truncated = 0;
};

// Payload decoders.  Absent fields keep protobuf defaults (0); unknown fields are skipped; return false on
// malformed input.
bool decodeObjectDirection(const uint8_t* p, size_t n, uint32_t& objectId, float& azimuthAngle, float& zenithAngle);
bool decodeObjectDistance(const uint8_t* p, size_t n, uint32_t& objectId, float& distance);
bool decodeObjectType(const uint8_t* p, size_t n, uint32_t& objectId, uint32_t& type);
bool decodeGeolocation(const uint8_t* p, size_t n, double& latitude, double& longitude, float& altitude, float& heading);
bool decodeGeodeticWgs84(const uint8_t* p, size_t n, double& latitude, double& longitude);
bool decodeGeodeticHeading(const uint8_t* p, size_t n, float& northHeading);
bool decodeAngularVelocity(const uint8_t* p, size_t n, float& x, float& y, float& z);

// Deterministic replay of the reference's data triggers (main: 71-106) + Slam::nextCone / nextPose /
// nextSplitPose / nextYawRate (slam.cpp:67-219) on recorded time: cone messages of sender `detectConeId` go
// through a FrameAssembler driven by the envelopes' sampleTimeStamp; pose messages of sender `estimationId`
// update the odometry (Geolocation: wgs84 -> Cartesian about the GPS reference + heading; split readings:
// position and heading separately, heading about PI) and the yaw rate (angularVelocityZ / 4).  Whenever the
// assembler releases a keyframe, `onFrame` gets what Slam::performSLAM would: the 4 x N frame and the
// odometry / yaw state at that moment.
struct ReplayConfig {
  uint32_t detectConeId = 116, estimationId = 112;   // reference example command line (main.cpp:55)
  int32_t gatheringTimeMs = 10;
  double timeBetweenKeyframes = 0.5;
  double refLatitude = 57.70924648, refLongitude = 11.9462;
};
struct ReplayFrame {
  slamtypes::MatrixXd cones;      // 4 x N
  double odometry[3];             // m_odometryData
  float yawRate;                  // m_yawRate
  double yawElapsed;              // |yaw stamp - last cone stamp| seconds (slam.cpp:309)
  int64_t time_us;                // recorded time the frame was released at
};
struct ReplayStats {
  size_t envelopes = 0, coneMessages = 0, poseMessages = 0, ignoredSender = 0, ignoredType = 0, malformed = 0;
  int framesGathered = 0, framesDroppedByKeyframeGate = 0, framesEmitted = 0;
};
ReplayStats replay(Reader& reader, const ReplayConfig& cfg, const std::function<void(const ReplayFrame&)>& onFrame);

}  // namespace slamrec

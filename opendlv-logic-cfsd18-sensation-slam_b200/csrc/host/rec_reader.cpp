// rec_reader.cpp -- see rec_reader.hpp.
#include "rec_reader.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>

#include "wgs84.hpp"

namespace slamrec {
namespace {

// ---- protobuf wire format, the subset cluon's ToProtoVisitor emits ----
enum WireType { VARINT = 0, EIGHT_BYTES = 1, LENGTH_DELIMITED = 2, FOUR_BYTES = 5 };

struct Cursor {
  const uint8_t* p;
  size_t n, off;
  bool ok;
  Cursor(const uint8_t* data, size_t len) : p(data), n(len), off(0), ok(true) {}
  bool done() const { return off >= n; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (off >= n) { ok = false; return 0; }
      const uint8_t b = p[off++];
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  float fixed32() {
    if (n - off < 4) { ok = false; return 0.0f; }
    uint32_t u = (uint32_t)p[off] | (uint32_t)p[off + 1] << 8 | (uint32_t)p[off + 2] << 16 | (uint32_t)p[off + 3] << 24;
    off += 4;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
  }
  double fixed64() {
    if (n - off < 8) { ok = false; return 0.0; }
    uint64_t u = 0;
    for (int k = 0; k < 8; k++) u |= (uint64_t)p[off + k] << (8 * k);
    off += 8;
    double d;
    std::memcpy(&d, &u, 8);
    return d;
  }
  Cursor sub() {  // length-delimited field
    const uint64_t len = varint();
    // len is an untrusted 64-bit varint: compare against the bytes LEFT (off <= n always holds), never
    // form off + len, which wraps for lengths near 2^64 and would pass the check
    if (!ok || len > (uint64_t)(n - off)) { ok = false; return Cursor(p, 0); }
    Cursor c(p + off, (size_t)len);
    off += (size_t)len;
    return c;
  }
  void skip(int wire) {
    if (wire == VARINT) varint();
    else if (wire == EIGHT_BYTES) { if (n - off < 8) ok = false; else off += 8; }
    else if (wire == FOUR_BYTES) { if (n - off < 4) ok = false; else off += 4; }
    else if (wire == LENGTH_DELIMITED) sub();
    else ok = false;
  }
};

inline int32_t zigzag32(uint64_t v) { return (int32_t)((uint32_t)(v >> 1) ^ (uint32_t)(-(int32_t)(v & 1))); }

// generic field walk: f(id, wire, cursor) consumes the value and returns true, or returns false to have it skipped
template <class F>
bool walk(Cursor& c, F f) {
  while (c.ok && !c.done()) {
    const uint64_t key = c.varint();
    if (!c.ok) break;
    const int wire = (int)(key & 7);
    const uint32_t id = (uint32_t)(key >> 3);
    if (!f(id, wire, c)) c.skip(wire);
  }
  return c.ok;
}

int64_t timestamp_us(Cursor c) {
  int32_t sec = 0, usec = 0;
  walk(c, [&](uint32_t id, int wire, Cursor& q) {
    if (wire != VARINT) return false;
    if (id == 1) { sec = zigzag32(q.varint()); return true; }
    if (id == 2) { usec = zigzag32(q.varint()); return true; }
    return false;
  });
  return (int64_t)sec * 1000000 + (int64_t)usec;  // cluon::time::toMicroseconds (4478-4480)
}

}  // namespace

bool Reader::open(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  m_file.assign(sz > 0 ? (size_t)sz : 0, 0);
  const size_t got = sz > 0 ? std::fread(m_file.data(), 1, (size_t)sz, f) : 0;
  std::fclose(f);
  if (got != m_file.size()) return false;
  attach(m_file.data(), m_file.size());
  return true;
}

void Reader::attach(const uint8_t* data, size_t len) {
  m_p = data;
  m_len = len;
  m_off = m_count = m_skipped = m_truncated = 0;
}

bool Reader::next(Envelope& e) {
  while (m_off + 5 <= m_len) {
    if (!(m_p[m_off] == 0x0D && m_p[m_off + 1] == 0xA4)) { m_off++; m_skipped++; continue; }
    const size_t len = (size_t)m_p[m_off + 2] | (size_t)m_p[m_off + 3] << 8 | (size_t)m_p[m_off + 4] << 16;
    if (m_off + 5 + len > m_len) { m_truncated = m_len - m_off; m_off = m_len; return false; }
    Cursor c(m_p + m_off + 5, len);
    Envelope out;
    const bool ok = walk(c, [&](uint32_t id, int wire, Cursor& q) {
      if (id == 1 && wire == VARINT) { out.dataType = zigzag32(q.varint()); return true; }
      if (id == 2 && wire == LENGTH_DELIMITED) { Cursor s = q.sub(); out.payload = s.p; out.payload_len = s.n; return true; }
      if (id == 3 && wire == LENGTH_DELIMITED) { out.sent_us = timestamp_us(q.sub()); return true; }
      if (id == 4 && wire == LENGTH_DELIMITED) { out.received_us = timestamp_us(q.sub()); return true; }
      if (id == 5 && wire == LENGTH_DELIMITED) { out.sample_us = timestamp_us(q.sub()); return true; }
      if (id == 6 && wire == VARINT) { out.senderStamp = (uint32_t)q.varint(); return true; }
      return false;
    });
    if (!ok) { m_off++; m_skipped++; continue; }  // a header look-alike inside garbage: resync
    m_off += 5 + len;
    m_count++;
    e = out;
    return true;
  }
  if (m_off < m_len) { m_truncated = m_len - m_off; m_off = m_len; }
  return false;
}

bool decodeObjectDirection(const uint8_t* p, size_t n, uint32_t& objectId, float& az, float& zen) {
  objectId = 0; az = 0; zen = 0;
  Cursor c(p, n);
  return walk(c, [&](uint32_t id, int wire, Cursor& q) {
    if (id == 1 && wire == VARINT) { objectId = (uint32_t)q.varint(); return true; }
    if (id == 2 && wire == FOUR_BYTES) { az = q.fixed32(); return true; }
    if (id == 3 && wire == FOUR_BYTES) { zen = q.fixed32(); return true; }
    return false;
  });
}

bool decodeObjectDistance(const uint8_t* p, size_t n, uint32_t& objectId, float& distance) {
  objectId = 0; distance = 0;
  Cursor c(p, n);
  return walk(c, [&](uint32_t id, int wire, Cursor& q) {
    if (id == 1 && wire == VARINT) { objectId = (uint32_t)q.varint(); return true; }
    if (id == 2 && wire == FOUR_BYTES) { distance = q.fixed32(); return true; }
    return false;
  });
}

bool decodeObjectType(const uint8_t* p, size_t n, uint32_t& objectId, uint32_t& type) {
  objectId = 0; type = 0;
  Cursor c(p, n);
  return walk(c, [&](uint32_t id, int wire, Cursor& q) {
    if (id == 1 && wire == VARINT) { objectId = (uint32_t)q.varint(); return true; }
    if (id == 2 && wire == VARINT) { type = (uint32_t)q.varint(); return true; }
    return false;
  });
}

bool decodeGeolocation(const uint8_t* p, size_t n, double& lat, double& lon, float& alt, float& heading) {
  lat = lon = 0; alt = heading = 0;
  Cursor c(p, n);
  return walk(c, [&](uint32_t id, int wire, Cursor& q) {
    if (id == 1 && wire == EIGHT_BYTES) { lat = q.fixed64(); return true; }
    if (id == 2 && wire == EIGHT_BYTES) { lon = q.fixed64(); return true; }
    if (id == 3 && wire == FOUR_BYTES) { alt = q.fixed32(); return true; }
    if (id == 4 && wire == FOUR_BYTES) { heading = q.fixed32(); return true; }
    return false;
  });
}

bool decodeGeodeticWgs84(const uint8_t* p, size_t n, double& lat, double& lon) {
  lat = lon = 0;
  Cursor c(p, n);
  return walk(c, [&](uint32_t id, int wire, Cursor& q) {
    if (id == 1 && wire == EIGHT_BYTES) { lat = q.fixed64(); return true; }
    if (id == 3 && wire == EIGHT_BYTES) { lon = q.fixed64(); return true; }  // longitude has id 3 (odvd:145-148)
    return false;
  });
}

bool decodeGeodeticHeading(const uint8_t* p, size_t n, float& northHeading) {
  northHeading = 0;
  Cursor c(p, n);
  return walk(c, [&](uint32_t id, int wire, Cursor& q) {
    if (id == 1 && wire == FOUR_BYTES) { northHeading = q.fixed32(); return true; }
    return false;
  });
}

bool decodeAngularVelocity(const uint8_t* p, size_t n, float& x, float& y, float& z) {
  x = y = z = 0;
  Cursor c(p, n);
  return walk(c, [&](uint32_t id, int wire, Cursor& q) {
    if (wire != FOUR_BYTES) return false;
    if (id == 1) { x = q.fixed32(); return true; }
    if (id == 2) { y = q.fixed32(); return true; }
    if (id == 3) { z = q.fixed32(); return true; }
    return false;
  });
}

ReplayStats replay(Reader& reader, const ReplayConfig& cfg, const std::function<void(const ReplayFrame&)>& onFrame) {
  ReplayStats st;
  FrameAssembler assembler(cfg.gatheringTimeMs, cfg.timeBetweenKeyframes);
  const double ref[2] = {cfg.refLatitude, cfg.refLongitude};
  double odometry[3] = {0, 0, 0};          // m_odometryData
  float yawRate = 0.0f;                    // m_yawRate
  int64_t yaw_us = 0, lastCone_us = 0;     // m_yawReceivedTime, m_lastTimeStamp
  auto release = [&](int64_t now_us) {
    ReplayFrame fr;
    if (!assembler.poll(now_us, fr.cones)) return;
    fr.odometry[0] = odometry[0]; fr.odometry[1] = odometry[1]; fr.odometry[2] = odometry[2];
    fr.yawRate = yawRate;
    fr.yawElapsed = std::fabs(static_cast<double>(yaw_us - lastCone_us)) / 1000000;  // slam.cpp:309
    fr.time_us = now_us;
    st.framesEmitted++;
    if (onFrame) onFrame(fr);
  };
  Envelope e;
  while (reader.next(e)) {
    st.envelopes++;
    const int64_t now = e.sample_us;
    // the reference's collector thread snapshots the frame gatheringTimeMs after its first message, whatever
    // arrives next; on recorded time that is "before the first envelope stamped later than the window"
    release(now);
    const bool cone = e.dataType == ID_OBJECT_DIRECTION || e.dataType == ID_OBJECT_DISTANCE || e.dataType == ID_OBJECT_TYPE;
    const bool pose = e.dataType == ID_GEOLOCATION || e.dataType == ID_GEODETIC_WGS84 || e.dataType == ID_GEODETIC_HEADING ||
                      e.dataType == ID_ANGULAR_VELOCITY;
    if (!cone && !pose) { st.ignoredType++; continue; }
    if (e.senderStamp != (cone ? cfg.detectConeId : cfg.estimationId)) { st.ignoredSender++; continue; }  // main: 71-97
    uint32_t id = 0, type = 0;
    float a = 0, b = 0, c3 = 0;
    double lat = 0, lon = 0;
    bool ok = true;
    switch (e.dataType) {
      case ID_OBJECT_DIRECTION:
        ok = decodeObjectDirection(e.payload, e.payload_len, id, a, b);
        if (ok) { lastCone_us = now; assembler.addDirection(id, a, b, now); st.coneMessages++; }
        break;
      case ID_OBJECT_DISTANCE:
        ok = decodeObjectDistance(e.payload, e.payload_len, id, a);
        if (ok) { lastCone_us = now; assembler.addDistance(id, a, now); st.coneMessages++; }
        break;
      case ID_OBJECT_TYPE:
        ok = decodeObjectType(e.payload, e.payload_len, id, type);
        if (ok) { lastCone_us = now; assembler.addType(id, type, now); st.coneMessages++; }
        break;
      case ID_GEOLOCATION: {  // Slam::nextPose, slam.cpp:185-210
        ok = decodeGeolocation(e.payload, e.payload_len, lat, lon, a, b);
        if (ok) {
          const double pos[2] = {lat, lon};
          double xy[2];
          slamwgs84::toCartesian(ref, pos, xy);
          odometry[0] = xy[0]; odometry[1] = xy[1]; odometry[2] = b;  // heading as sent (float widened)
          st.poseMessages++;
        }
        break;
      }
      case ID_GEODETIC_WGS84: {  // Slam::nextSplitPose, slam.cpp:153-175
        ok = decodeGeodeticWgs84(e.payload, e.payload_len, lat, lon);
        if (ok) {
          const double pos[2] = {lat, lon};
          double xy[2];
          slamwgs84::toCartesian(ref, pos, xy);
          odometry[0] = xy[0]; odometry[1] = xy[1];
          st.poseMessages++;
        }
        break;
      }
      case ID_GEODETIC_HEADING:  // slam.cpp:176-182
        ok = decodeGeodeticHeading(e.payload, e.payload_len, a);
        if (ok) { odometry[2] = slamwgs84::headingFromNorth(a); st.poseMessages++; }
        break;
      case ID_ANGULAR_VELOCITY:  // Slam::nextYawRate, slam.cpp:212-219
        ok = decodeAngularVelocity(e.payload, e.payload_len, a, b, c3);
        if (ok) { yawRate = c3 / 4; yaw_us = now; st.poseMessages++; }
        break;
      default: break;
    }
    if (!ok) st.malformed++;
  }
  // end of the recording: a frame still open is gathered as if the window had elapsed
  release(INT64_MAX / 4);
  st.framesGathered = assembler.framesGathered();
  st.framesDroppedByKeyframeGate = assembler.framesDroppedByKeyframeGate();
  return st;
}

}  // namespace slamrec

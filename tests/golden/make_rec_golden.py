"""Writes tests/golden/c1_head.rec with THE REFERENCE'S OWN cluon: the single-header library
/root/reference/src/cluon-complete-build.hpp, its message compiler and the message set generated from the
reference's .odvd (the three commands of the reference's CMakeLists.txt:55-69) serialise the first frames of
the C1 drive as the OD4 envelopes the perception and estimation services would send.  The product's reader
(csrc/host/rec_reader.cpp) has to get every field back (tests/test_rec_reader.py).  `messages(synth)` is
imported by the tests: the inputs are rebuilt from seeds, the .rec holds what cluon made of them."""
import os
import subprocess
import sys
import tempfile

import numpy as np

REF = os.environ.get("SLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, ROOT)

DETECT_CONE_ID, ESTIMATION_ID = 116, 112       # the reference's example command line (main.cpp:55)
REF_LAT, REF_LON = 57.70924648, 11.9462

WRITER = r'''
#include <cstdio>
#include <fstream>
#include <string>
#include "cluon-complete.hpp"
#include "opendlv-standard-message-set.hpp"
template <class M>
static void put(std::ofstream& out, M& msg, long long t_us, unsigned sender) {
  cluon::ToProtoVisitor v;
  msg.accept(v);
  cluon::data::TimeStamp ts;
  ts.seconds((int32_t)(t_us / 1000000)).microseconds((int32_t)(t_us % 1000000));
  cluon::data::Envelope e;
  e.dataType(M::ID()).serializedData(v.encodedData()).sent(ts).received(ts).sampleTimeStamp(ts).senderStamp(sender);
  const std::string s = cluon::serializeEnvelope(std::move(e));
  out.write(s.data(), (std::streamsize)s.size());
}
int main(int argc, char** argv) {
  std::ofstream out(argv[1], std::ios::binary);
  char k; long long t; unsigned sender; double a, b, c, d;
  while (std::scanf(" %c %lld %u %la %la %la %la", &k, &t, &sender, &a, &b, &c, &d) == 7) {
    if (k == 'D') { opendlv::logic::perception::ObjectDirection m; m.objectId((uint32_t)a).azimuthAngle((float)b).zenithAngle((float)c); put(out, m, t, sender); }
    else if (k == 'R') { opendlv::logic::perception::ObjectDistance m; m.objectId((uint32_t)a).distance((float)b); put(out, m, t, sender); }
    else if (k == 'T') { opendlv::logic::perception::ObjectType m; m.objectId((uint32_t)a).type((uint32_t)b); put(out, m, t, sender); }
    else if (k == 'G') { opendlv::logic::sensation::Geolocation m; m.latitude(a).longitude(b).altitude((float)c).heading((float)d); put(out, m, t, sender); }
    else if (k == 'W') { opendlv::proxy::GeodeticWgs84Reading m; m.latitude(a).longitude(b); put(out, m, t, sender); }
    else if (k == 'H') { opendlv::proxy::GeodeticHeadingReading m; m.northHeading((float)a); put(out, m, t, sender); }
    else if (k == 'Y') { opendlv::proxy::AngularVelocityReading m; m.angularVelocityX((float)a).angularVelocityY((float)b).angularVelocityZ((float)c); put(out, m, t, sender); }
    else if (k == 'X') { opendlv::proxy::GroundSpeedReading m; m.groundSpeed((float)a); put(out, m, t, sender); }  // a type Slam ignores
  }
  return 0;
}
'''


def messages(synth, n_frames=40):
    """(kind, t_us, sender, a, b, c, d) tuples in send order + the frames / poses they encode."""
    d = synth.trackdrive(1)
    msgs, frames, geo = [], [], []
    t0 = 1_530_000_000_000_000            # June 2018, microseconds
    rng = np.random.default_rng(7)
    for k in range(n_frames):
        t = t0 + k * 600_000               # a keyframe every 0.6 s of recorded time
        fr = np.asarray(d.frames[k], dtype=np.float64)
        p = d.poses_noisy[k]
        lat = REF_LAT + p[1] / 111_000.0
        lon = REF_LON + p[0] / 59_000.0
        if k % 5 == 4:                     # split pose every fifth frame, Geolocation otherwise
            msgs.append(("W", t - 3000, ESTIMATION_ID, lat, lon, 0.0, 0.0))
            msgs.append(("H", t - 2500, ESTIMATION_ID, float(np.float32(p[2] + 3.14159265)), 0.0, 0.0, 0.0))
        else:
            msgs.append(("G", t - 3000, ESTIMATION_ID, lat, lon, 12.5, float(np.float32(p[2]))))
        msgs.append(("Y", t - 2000, ESTIMATION_ID, 0.01, -0.02, float(np.float32(0.2 * np.sin(k))), 0.0))
        msgs.append(("G", t - 1500, ESTIMATION_ID + 1, 1.0, 2.0, 3.0, 4.0))          # wrong sender: ignored
        msgs.append(("X", t - 1000, ESTIMATION_ID, 7.5, 0.0, 0.0, 0.0))              # a type Slam does not consume
        n = fr.shape[1]
        order = rng.permutation(n)
        for j, i in enumerate(order):       # direction / distance / type of one object arrive interleaved
            msgs.append(("D", t + 30 * j, DETECT_CONE_ID, float(i), float(fr[0, i]), float(fr[1, i]), 0.0))
            msgs.append(("R", t + 30 * j + 10, DETECT_CONE_ID, float(i), float(fr[2, i]), 0.0, 0.0))
            msgs.append(("T", t + 30 * j + 20, DETECT_CONE_ID, float(i), float(fr[3, i]), 0.0, 0.0))
        msgs.append(("D", t + 5000, DETECT_CONE_ID + 1, 0.0, 45.0, 0.0, 0.0))        # wrong sender: ignored
        frames.append(fr)
        geo.append((lat, lon, float(p[2]), k % 5 == 4))
    return msgs, frames, geo


def main():
    if not os.path.exists(os.path.join(REF, "src", "cluon-complete-build.hpp")):
        sys.exit("reference tree not found")
    from conftest import load_pkg
    pkg = load_pkg()
    msgs, _, _ = messages(pkg.synth)
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink(os.path.join(REF, "src", "cluon-complete-build.hpp"), os.path.join(tmp, "cluon-complete.hpp"))
        os.symlink(os.path.join(tmp, "cluon-complete.hpp"), os.path.join(tmp, "cluon-complete.cpp"))
        base = ["g++", "-std=c++14", "-pthread", "-w", "-include", "linux/sockios.h"]
        subprocess.run(base + ["-o", os.path.join(tmp, "cluon-msc"), os.path.join(tmp, "cluon-complete.cpp"), "-D", "HAVE_CLUON_MSC"], check=True)
        odvd = [f for f in os.listdir(os.path.join(REF, "src")) if f.endswith(".odvd")][0]
        subprocess.run([os.path.join(tmp, "cluon-msc"), "--cpp-sources", "--cpp-add-include-file=opendlv-standard-message-set.hpp",
                        "--out=" + os.path.join(tmp, "opendlv-standard-message-set.cpp"), os.path.join(REF, "src", odvd)], check=True)
        subprocess.run([os.path.join(tmp, "cluon-msc"), "--cpp-headers", "--out=" + os.path.join(tmp, "opendlv-standard-message-set.hpp"),
                        os.path.join(REF, "src", odvd)], check=True)
        open(os.path.join(tmp, "writer.cpp"), "w").write(WRITER)
        subprocess.run(base + ["-O1", "-I" + tmp, os.path.join(tmp, "writer.cpp"), os.path.join(tmp, "opendlv-standard-message-set.cpp"),
                               "-o", os.path.join(tmp, "writer")], check=True)
        text = "\n".join("%s %d %d %s %s %s %s" % (m[0], m[1], m[2], float(m[3]).hex(), float(m[4]).hex(), float(m[5]).hex(), float(m[6]).hex())
                         for m in msgs) + "\n"
        out = os.path.join(HERE, "c1_head.rec")
        subprocess.run([os.path.join(tmp, "writer"), out], input=text, text=True, check=True)
    print("wrote %d envelopes, %d bytes" % (len(msgs), os.path.getsize(out)))


if __name__ == "__main__":
    main()

"""Generates tests/golden/cone_vectors.json by compiling THE REFERENCE'S OWN src/cone.cpp (from where it
lies; Eigen from the reference's thirdparty/) and calling Cone::getDirection / Cone::getDistance
(cone.cpp:34-53).  cone.hpp includes the cluon-generated "opendlv-standard-message-set.hpp", which needs
the reference's build system; a 20-line shim with the two message classes (fluent setter / getter
pairs, exactly the calls cone.cpp makes) stands in for it -- the arithmetic is the reference's.
Run in the build container; the JSON is committed (values as hex floats -> bit-for-bit comparison)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

REF = os.environ.get("SLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

SHIM = r'''
#pragma once
namespace opendlv { namespace logic { namespace perception {
class ObjectDirection {
 public:
  ObjectDirection& azimuthAngle(const float& v) { m_az = v; return *this; }
  float azimuthAngle() const { return m_az; }
  ObjectDirection& zenithAngle(const float& v) { m_zen = v; return *this; }
  float zenithAngle() const { return m_zen; }
 private:
  float m_az{0}, m_zen{0};
};
class ObjectDistance {
 public:
  ObjectDistance& distance(const float& v) { m_d = v; return *this; }
  float distance() const { return m_d; }
 private:
  float m_d{0};
};
}}}
'''

DRIVER = r'''
#include <cstdio>
#include "cone.hpp"
int main() {
  double cx, cy, px, py, pt;
  while (std::scanf(" %la %la %la %la %la", &cx, &cy, &px, &py, &pt) == 5) {
    Cone c(cx, cy, 1, 7);
    Eigen::Vector3d pose(px, py, pt);
    std::printf("%a %a %a\n", (double)c.getDirection(pose).azimuthAngle(), (double)c.getDirection(pose).zenithAngle(),
                (double)c.getDistance(pose).distance());
  }
  return 0;
}
'''


def cases():
    rng = np.random.default_rng(2018)
    out = []
    for _ in range(150):
        pose = (rng.uniform(-200, 200), rng.uniform(-200, 200), rng.uniform(-3.2, 3.2))
        r, a = rng.uniform(0.3, 60), rng.uniform(-np.pi, np.pi)
        out.append((pose[0] + r * np.cos(a), pose[1] + r * np.sin(a)) + pose)
    out += [(1.0, 0.0, 0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 0.0, 0.5), (-1.0, 0.0, 0.0, 0.0, 0.0), (-1.0, -0.0, 0.0, 0.0, 0.0),
            (0.0, 2.0, 0.0, 0.0, 1.5707963267948966), (3.0, 4.0, 0.0, 0.0, -3.1), (1e-9, 1e-9, 0.0, 0.0, 0.0)]
    return out


def main():
    if not os.path.exists(os.path.join(REF, "src", "cone.cpp")):
        sys.exit("reference tree not found")
    cs = cases()
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "opendlv-standard-message-set.hpp"), "w").write(SHIM)
        open(os.path.join(tmp, "drv.cpp"), "w").write(DRIVER)
        exe = os.path.join(tmp, "drv")
        subprocess.run(["g++", "-std=c++14", "-O2", "-w", "-I" + tmp, "-I" + os.path.join(REF, "src"), "-isystem",
                        os.path.join(REF, "thirdparty"), os.path.join(tmp, "drv.cpp"), os.path.join(REF, "src", "cone.cpp"),
                        "-o", exe], check=True)
        inp = "\n".join(" ".join(float(v).hex() for v in c) for c in cs) + "\n"
        lines = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split("\n")
    recs = []
    for c, line in zip(cs, lines):
        az, zen, dist = (float.fromhex(t) for t in line.split())
        recs.append({"cone": [float(c[0]).hex(), float(c[1]).hex()], "pose": [float(v).hex() for v in c[2:]],
                     "azimuth": az.hex(), "zenith": zen.hex(), "distance": dist.hex(),
                     "readable": {"cone": list(c[:2]), "pose": list(c[2:]), "azimuth": az, "distance": dist}})
    json.dump({"source": "reference src/cone.cpp (Cone::getDirection / getDistance) compiled with g++ -std=c++14 -O2 by "
                         "tests/golden/make_cone_golden.py; message classes shimmed", "n": len(recs), "vectors": recs},
              open(os.path.join(HERE, "cone_vectors.json"), "w"), indent=1)
    print("wrote %d vectors" % len(recs))


if __name__ == "__main__":
    main()

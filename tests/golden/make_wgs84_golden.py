"""Generates tests/golden/wgs84_vectors.json by compiling and running THE REFERENCE'S OWN header
/root/reference/src/WGS84toCartesian.hpp (never copied: included from where it lies) on a fixed set of
inputs.  Run in the build container (the GPU box has no /root/reference); the JSON is committed.
Values are stored as C99 hex floats, so the comparison in tests/test_wgs84.py is bit for bit."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

REF = os.environ.get("SLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

DRIVER = r'''
#include <cstdio>
#include <array>
#include "WGS84toCartesian.hpp"
int main() {
  char kind; double a, b, c, d;
  while (std::scanf(" %c %la %la %la %la", &kind, &a, &b, &c, &d) == 5) {
    std::array<double, 2> ref{a, b}, in{c, d};
    std::array<double, 2> out = (kind == 'T') ? wgs84::toCartesian(ref, in) : wgs84::fromCartesian(ref, in);
    std::printf("%a %a\n", out[0], out[1]);
  }
  return 0;
}
'''


def cases():
    rng = np.random.default_rng(1884)
    refs = [(57.70924648, 11.9462),      # the reference's example command line (refLatitude / refLongitude)
            (0.0, 0.0), (-33.8688, 151.2093), (48.1351, 11.5820), (89.9, 45.0), (37.7749, -122.4194)]
    to, fr = [], []
    # the survey's known answers
    to.append((refs[0], (57.7095, 11.9470)))
    to.append((refs[0], refs[0]))
    for ref in refs:
        for _ in range(12):   # within the +-200 m the path accepts (slam.cpp:300) and a little beyond
            to.append((ref, (ref[0] + rng.uniform(-4e-3, 4e-3), ref[1] + rng.uniform(-8e-3, 8e-3))))
        for _ in range(3):    # far away
            to.append((ref, (ref[0] + rng.uniform(-1, 1) * min(1.0, 89.99 - abs(ref[0])), ref[1] + rng.uniform(-3, 3))))
    # branches: latitude ~ 0, the poles, beyond the pole, longitude beyond 10 rad
    to += [(refs[1], (0.0, 0.01)), (refs[1], (5e-9, 0.01)), (refs[0], (90.0, 11.0)), (refs[0], (-90.0, 11.0)),
           (refs[0], (90.0000001, 11.0)), (refs[0], (57.7, 600.0)), (refs[2], (-33.87, 151.21))]
    for ref in (refs[0], refs[2], refs[3], refs[5]):
        for _ in range(8):
            fr.append((ref, (rng.uniform(-200, 200), rng.uniform(-200, 200))))
        fr.append((ref, (0.0, 0.0)))
        fr.append((ref, (rng.uniform(-1500, 1500), rng.uniform(-1500, 1500))))
    fr.append((refs[0], (47.688789808649354, 28.235440887535205)))   # survey: back to (57.70950648000008, 11.946999999999969)
    return to, fr


def main():
    hdr = os.path.join(REF, "src", "WGS84toCartesian.hpp")
    if not os.path.exists(hdr):
        sys.exit("reference header not found: " + hdr)
    to, fr = cases()
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "drv.cpp")
        open(src, "w").write(DRIVER)
        exe = os.path.join(tmp, "drv")
        # the reference's own flags (CMakeLists.txt:35-38: C++14, -O2, x86-64 baseline)
        subprocess.run(["g++", "-std=c++14", "-O2", "-I" + os.path.join(REF, "src"), src, "-o", exe], check=True)
        lines = ["T %s %s %s %s" % tuple(float(v).hex() for v in (r[0], r[1], p[0], p[1])) for r, p in to]
        lines += ["F %s %s %s %s" % tuple(float(v).hex() for v in (r[0], r[1], p[0], p[1])) for r, p in fr]
        out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    recs = []
    for (kind, (r, p)), line in zip([("to", c) for c in to] + [("from", c) for c in fr], out):
        a, b = line.split()
        recs.append({"kind": kind, "ref": [float(r[0]).hex(), float(r[1]).hex()], "in": [float(p[0]).hex(), float(p[1]).hex()],
                     "out": [float.fromhex(a).hex(), float.fromhex(b).hex()],
                     "readable": {"ref": list(r), "in": list(p), "out": [float.fromhex(a), float.fromhex(b)]}})
    doc = {"source": "reference src/WGS84toCartesian.hpp compiled with g++ -std=c++14 -O2 by tests/golden/make_wgs84_golden.py",
           "n": len(recs), "vectors": recs}
    json.dump(doc, open(os.path.join(HERE, "wgs84_vectors.json"), "w"), indent=1)
    print("wrote %d vectors" % len(recs))


if __name__ == "__main__":
    main()

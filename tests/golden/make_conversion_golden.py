"""tests/golden/conversion_vectors.json: the reference's private conversion helpers transformConeToCoG
(slam.cpp:513-523), Spherical2Cartesian (637-654) and coneToGlobal (499-510), called in the reference's REAL
slam.cpp (oracle/_ref/ref_slam_replay --conv) on azimuths over the whole circle (0 -> NaN, +-180, beyond),
non-zero zeniths, tiny and huge ranges and random poses.  Hex floats; the oracle must match bit for bit."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def cases():
    rng = np.random.default_rng(637)
    out = []
    for _ in range(200):
        out.append((rng.uniform(-100, 100), rng.choice([0.0, 0.0, rng.uniform(-20, 20)]), rng.uniform(0.3, 60), float(rng.integers(1, 5)),
                    rng.uniform(-200, 200), rng.uniform(-200, 200), rng.uniform(-3.2, 3.2)))
    for az in (0.0, -0.0, 1e-3, -1e-3, 1e-12, 90.0, -90.0, 179.999, 180.0, -180.0, 181.0, 270.0, 360.0, -725.5):
        for zen in (0.0, 1.0, -45.0, 90.0):
            for r in (1e-6, 1.5, 5.0, 1e4):
                out.append((az, zen, r, 2.0, 3.0, -2.0, 0.7))
    out = [tuple(float(np.float32(v)) if i < 3 else float(v) for i, v in enumerate(c)) for c in out]   # wire fields are float32
    return out


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_slam_replay")
    subprocess.run(["sh", os.path.join(ROOT, "oracle", "build_ref_slam.sh")], check=True)
    cs = cases()
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.txt"), os.path.join(tmp, "out.txt")
        open(fin, "w").write("\n".join(" ".join(float(v).hex() for v in c) for c in cs) + "\n")
        subprocess.run([exe, "--conv", fin, fout], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
        lines = open(fout).read().split("\n")
    recs = []
    for c, ln in zip(cs, lines):
        v = [float.fromhex(t) for t in ln.split()]
        recs.append({"in": [float(x).hex() for x in c], "cog": [float(x).hex() for x in v[0:2]], "xyz": [float(x).hex() for x in v[2:5]],
                     "global": [float(x).hex() for x in v[5:8]]})
    json.dump({"source": "reference src/slam.cpp private helpers via oracle/_ref/ref_slam_replay --conv", "n": len(recs), "vectors": recs},
              open(os.path.join(HERE, "conversion_vectors.json"), "w"), indent=0)
    print("wrote %d vectors" % len(recs))


if __name__ == "__main__":
    main()

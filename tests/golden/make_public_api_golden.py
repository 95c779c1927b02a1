"""Drives THE REFERENCE'S OWN class Slam through its PUBLIC interface (constructor with an OD4 session, nextPose /
nextYawRate / nextCone fed with cluon Envelopes, draw* to read the result; integration/public_api_harness.cpp built
against the unmodified src/slam.cpp by integration/build.sh) and writes tests/golden/public_api_reference.npz:

  <drive>_map        final map: x, y, type, id per cone (drawCones)
  <drive>_poses      the poses performSLAM stored (drawPoses)
  <drive>_graph_ptr / _graph_ids   drawGraph(): per stored pose the map cones its frame was associated with / created
  <drive>_current    drawCurrentPose()
  <drive>_frames     per input frame: number of stored poses and graph rows after it (which frames were dropped)

The same envelopes go through the PATCHED tree over the CUDA back end in tests/test_integration_gpu.py.  Drives:
  c1     configuration C1 (1,000 frames, loop closure, 25 localiser frames)
  odd    a noisy two-lap loop of 320 frames with a yaw rate on both sides of the (0, 1 s) window of the heading
         correction, absent objectIds (zero columns -> NaN cones), non-integer and large cone types, ranges beyond the
         mapping threshold, frames behind the 200 m gate
Run in the build container; the .npz and the drive files are committed."""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, ROOT)
from conftest import load_pkg  # noqa: E402


def write_drive(path, frames, poses, thr, map_thr, yaw):
    with open(path, "wb") as f:
        f.write(struct.pack("<idd", -len(frames), thr, map_thr))
        for k, (fr, p) in enumerate(zip(frames, poses)):
            fr = np.asfortranarray(fr, dtype=np.float64).reshape(4, -1, order="F")
            f.write(struct.pack("<fq", float(yaw[k][0]), int(yaw[k][1])))
            f.write(np.asarray(p, dtype=np.float64).tobytes())
            f.write(struct.pack("<i", fr.shape[1]))
            f.write(fr.tobytes(order="F"))


def parse(path):
    fx = float.fromhex
    M, P, G, F, C, D, S = [], [], [], [], None, 0, 0
    for ln in open(path).read().split("\n"):
        t = ln.split()
        if not t:
            continue
        if t[0] == "M":
            M.append((fx(t[2]), fx(t[3]), float(t[4]), float(t[5])))
        elif t[0] == "P":
            P.append([fx(v) for v in t[2:5]])
        elif t[0] == "G":
            G.append([int(v) for v in t[3:3 + int(t[2])]])
        elif t[0] == "F":
            F.append([int(t[2]), int(t[3])])
        elif t[0] == "C":
            C = [fx(v) for v in t[1:4]]
        elif t[0] == "D":
            D = int(t[1])
        elif t[0] == "S":
            S = int(t[1])
    ptr = np.zeros(len(G) + 1, dtype=np.int64)
    ids = []
    for k, r in enumerate(G):
        ids += r
        ptr[k + 1] = len(ids)
    return dict(map=np.array(M, dtype=np.float64).reshape(-1, 4), poses=np.array(P, dtype=np.float64).reshape(-1, 3),
                graph_ptr=ptr, graph_ids=np.array(ids, dtype=np.int64), current=np.array(C, dtype=np.float64),
                frames=np.array(F, dtype=np.int64).reshape(-1, 2), dropped=np.array([D])), S


def run(exe, drive, gathering_ms=10, timeout=900, attempts=4):
    """One replay; repeated while the harness reports that a frame's messages may have straddled a gathering window
    (its thread lost the CPU in the middle of a frame: the reference's collector runs on the wall clock)."""
    for _ in range(attempts):
        with tempfile.TemporaryDirectory() as tmp:
            out = os.path.join(tmp, "out.txt")
            subprocess.run([exe, drive, out, str(gathering_ms)], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=timeout)
            r, suspect = parse(out)
        if suspect == 0:
            return r
    raise RuntimeError("every attempt had a frame whose messages took more than half a gathering window to send")


def odd_drive(synth):
    rng = np.random.default_rng(77)
    trk = synth.ellipse_track(n_pairs=26, a=22.0, b=11.0, half_width=1.5)
    d = synth.simulate_drive(trk, 320, s_step=1.6 * trk.length / 320, seed=45, sigma_xy=0.06, sigma_th=0.006, closed=True)
    frames = [np.array(f, dtype=np.float64, order="F").reshape(4, -1, order="F").copy() for f in d.frames]
    poses = np.array(d.poses_noisy, dtype=np.float64).copy()
    yaw = []
    for k, fr in enumerate(frames):
        if fr.shape[1] > 2 and k % 7 == 3:      # an absent objectId in the middle of the frame: zero column
            fr[:, 1] = 0.0
        if fr.shape[1] > 1 and k % 11 == 5:     # non-integer / large types (wire type is uint32: truncated by the sender)
            fr[3, 0] = 4.0
        if fr.shape[1] > 1 and k % 13 == 6:     # beyond the mapping threshold
            fr[2, -1] = np.float32(60.0 + k * 0.01)
        if k in (40, 41, 200):                  # behind the 200 m gate of performSLAM
            poses[k, 0] = 250.0
        y = np.float32(rng.normal(0, 0.2))
        el = int(rng.choice([0, 20000, 400000, 990000, 1000000, 1500000, -30000]))
        yaw.append((y, el))
    return frames, poses, yaw


def main():
    subprocess.run(["sh", os.path.join(ROOT, "integration", "build.sh")], check=True)
    exe = os.path.join(ROOT, "integration", "_build", "ref_public_replay")
    synth = load_pkg().synth
    out = {}
    d = synth.trackdrive(1)
    yaw0 = [(0.0, 500 * 1000000)] * len(d.frames)   # far outside the (0, 1 s) window: no heading correction
    drives = {"c1": (d.frames, d.poses_noisy, yaw0)}
    drives["odd"] = odd_drive(synth)
    # six of the adversarial replays that pin the back half directly (make_fuzz_reference_replay.py), now through the
    # Envelopes: other gates (sameConeThreshold 0.8 / 2.0, coneMappingThreshold 9), repeated and scrambled columns,
    # non-integer types (truncated by the uint32 wire field), NaN cones in open drives, random yaw rates
    from make_fuzz_reference_replay import scenarios, scenarios_yaw
    extra = {}
    far = (0.0, 500 * 1000000)
    for name, frames, poses, thr, map_thr in scenarios(synth):
        if name in ("loop1", "loop2", "loop5", "nan0", "nan2"):
            extra[name] = (frames, poses, [far] * len(frames), thr, map_thr)
    for name, frames, poses, thr, map_thr, yaw in scenarios_yaw(synth):
        if name == "yaw1":
            extra[name] = (frames, poses, yaw, thr, map_thr)
    drives = {k: v + (synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD) for k, v in drives.items()}
    drives.update(extra)
    for name, (frames, poses, yaw, thr, map_thr) in drives.items():
        path = os.path.join(HERE, "public_api_drive_%s.bin" % name)
        write_drive(path, frames, poses, thr, map_thr, yaw)
        r = run(exe, path)
        r2 = run(exe, path)   # the front half runs on the wall clock: the result must not depend on it
        for k in r:
            assert np.array_equal(r[k], r2[k], equal_nan=True), "reference run not reproducible: %s %s" % (name, k)
            out["%s_%s" % (name, k)] = r[k]
        print(name, "frames", len(frames), "stored poses", len(r["poses"]), "map", len(r["map"]), "graph entries",
              len(r["graph_ids"]), "dropped", int(r["dropped"][0]))
    np.savez_compressed(os.path.join(HERE, "public_api_reference.npz"), **out)


if __name__ == "__main__":
    main()

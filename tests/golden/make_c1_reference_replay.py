"""Replays configuration C1 (trackdrive loop, seed 18) through THE REFERENCE'S OWN src/slam.cpp, compiled
from where it lies by oracle/build_ref_slam.sh (g2o replaced by the facade over the oracle's restated
Gauss-Newton), and writes tests/golden/c1_replay_reference.npz:

  frame_map_size, frame_cci, frame_loop_closing, frame_loop_closed, frame_pose_id  per frame
  row_ptr, row_ids        the row performSLAM appended to m_connectivityGraph per frame = the map cones the
                          frame's observations were associated with (or created as), in observation order
  map_x, map_y, map_type, map_id   final map (after the optimise burst at loop closure)
  poses                   the poses performSLAM stored (m_poses)
  vertices                the pose vertices of the graph at the end (optimised)
  chi2                    chi2 per iteration of the last optimise call

Everything up to the optimise call is computed by reference code alone, so the association rows, map sizes,
current-cone indices and the loop-closure frame pin the oracle's restatement of SURVEY 8(a) rows 1-7 and 19
(tests/test_pinned_by_reference.py).  Run in the build container; the .npz is committed."""
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


def replay(frames, poses, thr, map_thr, exe=None, yaw=None):
    """yaw: optional per-frame (float32 yawRate, int elapsed_us) -> extended records (heading correction)."""
    exe = exe or os.path.join(ROOT, "oracle", "_ref", "ref_slam_replay")
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "frames.bin"), os.path.join(tmp, "out.txt")
        with open(fin, "wb") as f:
            f.write(struct.pack("<idd", -len(frames) if yaw is not None else len(frames), thr, map_thr))
            for k, (fr, p) in enumerate(zip(frames, poses)):
                fr = np.asfortranarray(fr, dtype=np.float64)
                if yaw is not None:
                    f.write(struct.pack("<fq", float(yaw[k][0]), int(yaw[k][1])))
                f.write(np.asarray(p, dtype=np.float64).tobytes())
                f.write(struct.pack("<i", fr.shape[1]))
                f.write(fr.tobytes(order="F"))
        subprocess.run([exe, fin, fout], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
        lines = open(fout).read().split("\n")
    fx = lambda t: float.fromhex(t)
    F, rows, M, P, V, chi2 = [], [], [], [], [], []
    timing = None
    for ln in lines:
        t = ln.split()
        if not t:
            continue
        if t[0] == "F":
            F.append([int(v) for v in t[2:7]])
            n = int(t[7])
            rows.append([int(v) for v in t[8:8 + n]] if n >= 0 else None)
        elif t[0] == "M":
            M.append((fx(t[2]), fx(t[3]), int(t[4]), int(t[5])))
        elif t[0] == "P":
            P.append([fx(v) for v in t[2:5]])
        elif t[0] == "V":
            V.append([fx(v) for v in t[2:5]])
        elif t[0] == "C":
            chi2.append(fx(t[1]))
        elif t[0] == "T":
            timing = dict(mapping_frames=int(t[1]), mapping_s=float(t[2]), closing_frames=int(t[3]), closing_s=float(t[4]),
                          localise_frames=int(t[5]), localise_s=float(t[6]))
    F = np.array(F, dtype=np.int64).reshape(-1, 5)
    row_ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    ids = []
    for k, r in enumerate(rows):
        ids += (r or [])
        row_ptr[k + 1] = len(ids)
    replay.last_timing = timing   # not part of the fixture (machine dependent)
    return dict(frame_map_size=F[:, 0], frame_cci=F[:, 1], frame_loop_closing=F[:, 2], frame_loop_closed=F[:, 3],
                frame_pose_id=F[:, 4], row_ptr=row_ptr, row_ids=np.array(ids, dtype=np.int64),
                row_present=np.array([r is not None for r in rows]),
                map_x=np.array([m[0] for m in M]), map_y=np.array([m[1] for m in M]),
                map_type=np.array([m[2] for m in M], dtype=np.int64), map_id=np.array([m[3] for m in M], dtype=np.int64),
                poses=np.array(P).reshape(-1, 3), vertices=np.array(V).reshape(-1, 3), chi2=np.array(chi2))


def main():
    subprocess.run(["sh", os.path.join(ROOT, "oracle", "build_ref_slam.sh")], check=True)
    pkg = load_pkg()
    synth = pkg.synth
    d = synth.trackdrive(1)
    r = replay(d.frames, d.poses_noisy, synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    np.savez_compressed(os.path.join(HERE, "c1_replay_reference.npz"), **r)
    closed = int(np.argmax(r["frame_loop_closed"])) if r["frame_loop_closed"].any() else -1
    print("frames %d, map %d cones, loop closed at frame %d, %d association entries, chi2 %s"
          % (len(r["frame_map_size"]), len(r["map_x"]), closed, len(r["row_ids"]), r["chi2"][-3:] if len(r["chi2"]) else None))


if __name__ == "__main__":
    main()

"""Differential replays against THE REFERENCE'S OWN src/slam.cpp (oracle/_ref/ref_slam_replay, see
make_c1_reference_replay.py): small noisy loops that close early and then localise, with adversarial
frames mixed in -- a column repeated inside a frame, a non-integer cone type, a range beyond the mapping
threshold, an empty frame, columns in scrambled order -- and open drives with azimuth-0 columns (NaN
cones, slam.cpp:515).  `scenarios(synth)` is imported by the tests so inputs are rebuilt from seeds and
only the reference's outputs are stored (tests/golden/fuzz_replay_reference.npz)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, ROOT)


def scenarios(synth):
    out = []
    for k in range(8):          # A: closed loops, loop closure + localiser phase
        rng = np.random.default_rng(100 + k)
        trk = synth.ellipse_track(n_pairs=22 + 2 * k, a=18.0 + 2 * k, b=9.0 + k, half_width=1.5)
        n = 260
        d = synth.simulate_drive(trk, n, s_step=1.45 * trk.length / n, seed=40 + k, sigma_xy=0.05 + 0.03 * k,
                                 sigma_th=0.004 + 0.002 * k, sigma_r=0.03 + 0.04 * k, sigma_az=0.15 + 0.2 * k)
        frames = [np.asfortranarray(f, dtype=np.float64).copy() for f in d.frames]
        for i, f in enumerate(frames):
            if i == 0 or f.shape[1] == 0:
                continue
            if i % 17 == 3:                                   # a column repeated inside the frame
                f = np.concatenate([f, f[:, :1]], axis=1)
            if i % 23 == 5:                                   # non-integer type (never matches, always new)
                f[3, rng.integers(f.shape[1])] = 1.5
            if i % 29 == 7:                                   # beyond the mapping threshold: neither matched nor added
                f[2, rng.integers(f.shape[1])] = 60.0
            if i % 31 == 11:                                  # empty frame (the map is not empty any more)
                f = np.zeros((4, 0))
            if i % 13 == 2 and f.shape[1] > 1:                # scrambled column order
                f = f[:, rng.permutation(f.shape[1])]
            frames[i] = np.asfortranarray(f)
        thr = [0.8, 1.2, 2.0][k % 3]
        map_thr = [50.0, 9.0][k % 2]
        out.append(("loop%d" % k, frames, d.poses_noisy.copy(), thr, map_thr))
    for k in range(4):          # B: open drives with NaN cones, no loop closure
        rng = np.random.default_rng(200 + k)
        trk = synth.ellipse_track(n_pairs=60, a=60.0, b=30.0, half_width=1.5)
        d = synth.simulate_drive(trk, 70, s_step=0.5, seed=70 + k, sigma_r=0.05, sigma_az=0.3)
        frames = [np.asfortranarray(f, dtype=np.float64).copy() for f in d.frames]
        for i, f in enumerate(frames):
            if i > 0 and i % 9 == 4 and f.shape[1] > 0:
                f[0, rng.integers(f.shape[1])] = 0.0          # azimuth 0 -> NaN cone
        out.append(("nan%d" % k, frames, d.poses_noisy.copy(), 1.2, 50.0))
    return out


def scenarios_yaw(synth):
    """Replays that exercise the heading correction of performSLAM (slam.cpp:309-318): a yaw rate and the time
    between the yaw reading and the last cone message per frame -- inside (0, 1) s the heading is corrected,
    at 0 or beyond 1 s it is not.  Returns (name, frames, poses, thr, map_thr, yaw[(float32 rate, elapsed_us)])."""
    out = []
    for k in range(4):
        rng = np.random.default_rng(300 + k)
        trk = synth.ellipse_track(n_pairs=24 + 2 * k, a=20.0 + 2 * k, b=10.0 + k, half_width=1.5)
        n = 240
        d = synth.simulate_drive(trk, n, s_step=1.4 * trk.length / n, seed=90 + k, sigma_r=0.05, sigma_az=0.3)
        frames = [np.asfortranarray(f, dtype=np.float64).copy() for f in d.frames]
        yaw = []
        for i in range(n):
            rate = np.float32(rng.normal(0.0, 0.3))
            el = int(rng.choice([0, 1, 999_999, 1_000_000, 1_000_001, 2_500_000, int(rng.integers(1, 1_000_000)), int(rng.integers(1, 1_000_000))]))
            yaw.append((rate, el))
        out.append(("yaw%d" % k, frames, d.poses_noisy.copy(), 1.2, 50.0, yaw))
    return out


def scenarios_gate(synth):
    """Replays shifted so that part of the loop lies beyond the 200 m sanity gate of performSLAM (slam.cpp:300-303:
    the frame is dropped before a pose is added), with big-orange (4) and negative cone types, zero ranges,
    azimuths beyond the half circle, zenith != 0 and single-column frames (which the localiser skips, slam.cpp:332,
    while the mapping phase does not) mixed in.  One drive never closes its loop, two close and localise."""
    out = []
    for k in range(3):
        rng = np.random.default_rng(400 + k)
        trk = synth.ellipse_track(n_pairs=24 + 2 * k, a=20.0 + 2 * k, b=10.0 + k, half_width=1.5)
        n = 300
        d = synth.simulate_drive(trk, n, s_step=1.5 * trk.length / n, seed=120 + k, sigma_r=0.05, sigma_az=0.3)
        poses = d.poses_noisy.copy()
        shift = [(188.0, 0.0), (0.0, -193.0), (-186.0, 190.0)][k]
        poses[:, 0] += shift[0]
        poses[:, 1] += shift[1]
        frames = [np.asfortranarray(f, dtype=np.float64).copy() for f in d.frames]
        for i, f in enumerate(frames):
            if i == 0 or f.shape[1] == 0:
                continue
            if i % 19 == 4:
                f[3, rng.integers(f.shape[1])] = 4.0                                  # big orange
            if i % 21 == 6:
                f[3, rng.integers(f.shape[1])] = -1.0                                 # negative type
            if i % 27 == 8:
                f[2, rng.integers(f.shape[1])] = 0.0                                  # zero range
            if i % 33 == 10:
                f[0, rng.integers(f.shape[1])] = [181.0, -270.0, 359.5][k]            # azimuth beyond the half circle
            if i % 7 == 3:
                f = f[:, :1]                                                          # single column
            if i % 37 == 12:
                f[1, rng.integers(f.shape[1])] = 5.0                                  # zenith != 0
            frames[i] = np.asfortranarray(f)
        out.append(("gate%d" % k, frames, poses, 1.2, 50.0))
    return out


def main():
    import subprocess
    from conftest import load_pkg
    from make_c1_reference_replay import replay
    subprocess.run(["sh", os.path.join(ROOT, "oracle", "build_ref_slam.sh")], check=True)
    pkg = load_pkg()
    store = {}
    for name, frames, poses, thr, map_thr in scenarios(pkg.synth):
        r = replay(frames, poses, thr, map_thr)
        for key, val in r.items():
            store[name + "/" + key] = val
        closed = int(np.argmax(r["frame_loop_closed"])) if r["frame_loop_closed"].any() else -1
        print("%-6s frames %3d  map %3d cones  loop closed at %4d  %5d association entries  NaN cones %d"
              % (name, len(frames), len(r["map_x"]), closed, len(r["row_ids"]), int(np.isnan(r["map_x"]).sum())))
    np.savez_compressed(os.path.join(HERE, "fuzz_replay_reference.npz"), **store)
    store = {}
    for name, frames, poses, thr, map_thr, yaw in scenarios_yaw(pkg.synth):
        r = replay(frames, poses, thr, map_thr, yaw=yaw)
        for key, val in r.items():
            store[name + "/" + key] = val
        closed = int(np.argmax(r["frame_loop_closed"])) if r["frame_loop_closed"].any() else -1
        print("%-6s frames %3d  map %3d cones  loop closed at %4d  %5d association entries"
              % (name, len(frames), len(r["map_x"]), closed, len(r["row_ids"])))
    np.savez_compressed(os.path.join(HERE, "fuzz_yaw_replay_reference.npz"), **store)

    store = {}
    for name, frames, poses, thr, map_thr in scenarios_gate(pkg.synth):
        r = replay(frames, poses, thr, map_thr)
        for key, val in r.items():
            store[name + "/" + key] = val
        closed = int(np.argmax(r["frame_loop_closed"])) if r["frame_loop_closed"].any() else -1
        print("%-6s frames %3d  map %3d cones  loop closed at %4d  %5d association entries  %3d frames rejected by the 200 m gate"
              % (name, len(frames), len(r["map_x"]), closed, len(r["row_ids"]), int((~r["row_present"]).sum())))
    np.savez_compressed(os.path.join(HERE, "fuzz_gate_replay_reference.npz"), **store)


if __name__ == "__main__":
    main()

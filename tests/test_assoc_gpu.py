"""Association on the B200 through the C ABI vs the CPU oracle: indices and status codes bit-exact
(north_star), coordinates to a few ulp (device sin/cos/asin vs glibc)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

THR, MAPTHR = 1.2, 50.0


def _ulp_close(a, b, ulps=16):
    a = np.asarray(a); b = np.asarray(b)
    both_nan = np.isnan(a) & np.isnan(b)
    # a few ulp of the value, plus a few ulp of the ~10 m operands a small result was cancelled from
    tol = ulps * np.spacing(np.maximum(np.abs(a), np.abs(b))) + 1e-13
    return np.all(both_nan | (np.abs(a - b) <= tol))


def test_conversion_matches_oracle(ctx, orc, c1_drive):
    fr = np.concatenate(c1_drive.frames[:40], axis=1)
    fr = np.asfortranarray(fr)
    fr[1, ::3] = 1.5  # non-zero zenith on some columns
    pose = np.array([3.0, -2.0, 0.7])
    g, l = ctx.cones_to_global(fr, pose)
    for i in range(fr.shape[1]):
        go = orc.cone_to_global(pose, fr[:, i])
        lo = orc.spherical2cartesian(fr[0, i], fr[1, i], fr[2, i])
        assert _ulp_close(g[i, :2], go[:2]) and g[i, 2] == go[2]
        assert _ulp_close(l[i], lo)


def test_conversion_vectors_from_the_reference_slam_cpp(ctx):
    """tests/golden/conversion_vectors.json (the reference's real transformConeToCoG / Spherical2Cartesian /
    coneToGlobal, make_conversion_golden.py) through slam_b200_cones_to_global: whole azimuth circle incl. 0 -> NaN
    and +-180 degrees, zenith != 0, ranges 1e-6 .. 1e4 m, at the 16-ulp bar (scaled with the range for the
    cancellation term), NaN in the same slots."""
    import json
    import os
    doc = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "conversion_vectors.json")))
    fx = float.fromhex
    by_pose = {}
    for v in doc["vectors"]:
        a = [fx(x) for x in v["in"]]
        by_pose.setdefault(tuple(a[4:7]), []).append((a[:4], [fx(x) for x in v["xyz"]], [fx(x) for x in v["global"]]))
    checked = 0
    for pose, rows in by_pose.items():
        fr = np.asfortranarray(np.array([r[0] for r in rows]).T)
        g, l = ctx.cones_to_global(fr, np.array(pose))
        for i, (inp, xyz, glob) in enumerate(rows):
            scale = max(1.0, abs(inp[2])) * 1e-13
            for got, want in ((l[i], np.array(xyz)), (g[i, :2], np.array(glob[:2]))):
                both_nan = np.isnan(got) & np.isnan(want)
                tol = 16 * np.spacing(np.maximum(np.abs(got), np.abs(want))) + scale
                assert np.all(both_nan | (np.abs(got - want) <= tol)), (inp, got, want)
            assert g[i, 2] == glob[2]
            checked += 1
    assert checked == doc["n"]


def test_conversion_nan_at_zero_azimuth(ctx):
    fr = np.asfortranarray(np.array([[0.0], [0.0], [5.0], [1.0]]))
    g, l = ctx.cones_to_global(fr, np.zeros(3))
    assert np.isnan(g[0, 0]) and np.isnan(g[0, 1]) and l[0, 2] == 0.0  # slam.cpp:515


def _lockstep_mapping(ctx, orc, frames, poses, thr=THR, map_thr=MAPTHR, stop_at_closure=True):
    """Drives the device map and an oracle map through the same frames; asserts bit-exact records."""
    ctx.map_clear()
    cap = sum(f.shape[1] for f in frames) + 8
    mx = np.zeros(cap); my = np.zeros(cap); mt = np.zeros(cap, dtype=np.int32)
    M = 0; cci = 0; lc = 0
    dcci = 0; dlc = 0
    for k, (fr, p) in enumerate(zip(frames, poses)):
        o = orc.assoc_map_frame(fr, p, thr, map_thr, mx, my, mt, M, cci, lc)
        d = ctx.assoc_map_frame(fr, p, thr, map_thr, dcci, dlc)
        assert np.array_equal(d["idx"], o["idx"]), f"frame {k}"
        assert np.array_equal(d["status"], o["status"]), f"frame {k}"
        assert (d["first"], d["lc_obs"], d["M"], d["cci"], d["loop_closing"]) == \
               (o["first"], o["lc_obs"], o["M"], o["cci"], o["loop_closing"]), f"frame {k}"
        assert _ulp_close(d["z"], o["z"]) and _ulp_close(d["g"], o["g"])
        M, cci, lc = o["M"], o["cci"], o["loop_closing"]
        dcci, dlc = d["cci"], d["loop_closing"]
        if lc and stop_at_closure:
            break
    dx, dy, dt = ctx.map_read()
    assert len(dx) == M and np.array_equal(dt, mt[:M])
    assert _ulp_close(dx, mx[:M]) and _ulp_close(dy, my[:M])
    return M, k


def test_mapping_phase_c1_replay_bit_exact(ctx, orc, c1_drive):
    M, k = _lockstep_mapping(ctx, orc, c1_drive.frames, c1_drive.poses_noisy)
    assert M == 300 and k > 900   # the whole lap, loop closure near the end


def test_mapping_phase_edge_cases(ctx, orc, synth):
    pose = np.array([0.0, 0.0, 0.1])
    frames = [
        np.zeros((4, 0), order="F"),                                              # empty frame
        np.asfortranarray(np.array([[10.0, 10.0, -20.0, 0.0, 30.0],               # two columns = the same new cone
                                    [0, 0, 0, 0, 0],                               # (second must match the first, created
                                    [5.0, 5.2, 7.0, 4.0, 60.0],                    #  in this very frame); az == 0 -> NaN cone;
                                    [1.0, 1.0, 2.0, 1.0, 2.0]])),                  # range 60 >= 50 -> not added
        np.asfortranarray(np.array([[10.0, -20.0, 12.0], [0, 0, 0], [5.0, 7.0, 5.1], [1.5, 2.0, 2.0]])),  # non-integer type
    ]
    poses = [pose, pose, pose + np.array([0.2, 0.0, 0.0])]
    _lockstep_mapping(ctx, orc, frames, poses, stop_at_closure=False)


def test_mapping_phase_mid_frame_loop_closure(ctx, orc, c1_drive):
    """An unsorted frame whose FIRST column closes the loop: every later column is skipped and
    the caller owes one optimise per remaining column (slam.cpp:625-633)."""
    frames = [f.copy(order="F") for f in c1_drive.frames]
    # find the closing frame with the oracle, then reverse its columns
    cap = 4000
    mx = np.zeros(cap); my = np.zeros(cap); mt = np.zeros(cap, dtype=np.int32)
    M = cci = lc = 0
    for k, (fr, p) in enumerate(zip(frames, c1_drive.poses_noisy)):
        o = orc.assoc_map_frame(fr, p, THR, MAPTHR, mx, my, mt, M, cci, lc)
        M, cci, lc = o["M"], o["cci"], o["loop_closing"]
        if lc:
            break
    frames[k] = np.asfortranarray(frames[k][:, ::-1])
    ctx.map_clear()
    mx[:] = 0; my[:] = 0; mt[:] = 0
    M = cci = lc = 0; dcci = dlc = 0
    for q in range(k + 1):
        o = orc.assoc_map_frame(frames[q], c1_drive.poses_noisy[q], THR, MAPTHR, mx, my, mt, M, cci, lc)
        d = ctx.assoc_map_frame(frames[q], c1_drive.poses_noisy[q], THR, MAPTHR, dcci, dlc)
        assert np.array_equal(d["idx"], o["idx"]) and np.array_equal(d["status"], o["status"])
        assert d["lc_obs"] == o["lc_obs"]
        M, cci, lc = o["M"], o["cci"], o["loop_closing"]; dcci, dlc = d["cci"], d["loop_closing"]
    assert o["lc_obs"] == 0 and np.all(o["status"][1:] == 3)


def test_mapping_phase_large_frame_many_tiles(ctx, orc, synth):
    """A frame and a map larger than one shared-memory tile (2048 cones) and than the warp count."""
    f = synth.cone_field(n_map=6000, n_obs=1500, seed=11, density=0.02)
    ctx.map_clear()
    ctx.map_append(f.map_x, f.map_y, f.map_type)
    cap = 6000 + 1500 + 1
    mx = np.zeros(cap); my = np.zeros(cap); mt = np.zeros(cap, dtype=np.int32)
    mx[:6000], my[:6000], mt[:6000] = f.map_x, f.map_y, f.map_type
    o = orc.assoc_map_frame(f.frame, f.pose, THR, 1e9, mx, my, mt, 6000, 0, 0)
    d = ctx.assoc_map_frame(f.frame, f.pose, THR, 1e9, 0, 0)
    assert np.array_equal(d["idx"], o["idx"]) and np.array_equal(d["status"], o["status"])
    assert d["M"] == o["M"] and d["cci"] == o["cci"] and d["lc_obs"] == o["lc_obs"]
    assert (o["status"] == 1).sum() > 50 and (o["status"] == 0).sum() > 1000


def test_localize_phase_matches_oracle(ctx, orc, synth, c1_drive):
    trk = c1_drive.track
    ctx.map_clear()
    ctx.map_append(trk.cones_xy[:, 0], trk.cones_xy[:, 1], trk.cones_type)
    mx, my, mt = trk.cones_xy[:, 0].copy(), trk.cones_xy[:, 1].copy(), trk.cones_type.copy()
    cci = dcci = 7
    for k in range(0, 1000, 37):
        fr = c1_drive.frames[k].copy(order="F")
        if k % 2:
            fr[3] = 3.0 - fr[3]          # swap types: exercises the asymmetric gate (slam.cpp:360)
        o = orc.assoc_localize_frame(fr, c1_drive.poses_noisy[k], THR, mx, my, mt, cci)
        d = ctx.assoc_localize_frame(fr, c1_drive.poses_noisy[k], THR, dcci)
        assert np.array_equal(d["idx"], o["idx"])
        assert (d["cci"], d["n_reobserved"]) == (o["cci"], o["n_reobserved"])
        if o["n_reobserved"] > 0:
            assert d["send_cone_data"] == o["send_cone_data"]
        cci, dcci = o["cci"], d["cci"]
    # nothing matched: current cone index is left alone (slam.cpp:387)
    far = np.asfortranarray(np.array([[5.0], [0.0], [3.0], [1.0]]))
    d = ctx.assoc_localize_frame(far, np.array([150.0, 150.0, 0.0]), THR, 11)
    assert d["n_reobserved"] == 0 and d["cci"] == 11 and d["idx"][0] == -1


@pytest.mark.parametrize("gate", [0, 1])
def test_bulk_match_only_vs_oracle(ctx, orc, synth, pkg, gate):
    f = synth.cone_field(n_map=200_000, n_obs=3000, seed=4)
    ctx.map_clear()
    ctx.map_append(f.map_x, f.map_y, f.map_type)
    o = orc.assoc_match_only(f.frame, f.pose, THR, gate, f.map_x, f.map_y, f.map_type)
    brute = ctx.assoc_bulk(f.frame, f.pose, THR, gate, pkg.capi.ALGO_BRUTE).copy()
    grid = ctx.assoc_bulk(f.frame, f.pose, THR, gate, pkg.capi.ALGO_GRID).copy()
    assert np.array_equal(brute, o["idx"])
    assert np.array_equal(grid, o["idx"])
    assert (o["idx"] >= 0).sum() > 2000
    # no decision sits within rounding distance of the threshold, so the few-ulp difference of the
    # device trig cannot flip one
    assert o["min_margin"] > 1e-9


def test_bulk_edge_cases(ctx, synth, pkg):
    f = synth.cone_field(n_map=5000, n_obs=400, seed=5)
    ctx.map_clear()
    assert np.all(ctx.assoc_bulk(f.frame, f.pose, THR, 0, pkg.capi.ALGO_GRID) == -1)   # empty map
    ctx.map_append(f.map_x, f.map_y, f.map_type)
    assert len(ctx.assoc_bulk(f.frame[:, :0], f.pose, THR, 0, pkg.capi.ALGO_GRID)) == 0  # empty batch
    fr = f.frame.copy(order="F")
    fr[0, 5] = 0.0                                    # NaN observation never matches
    fr[2, 6] = 1e7                                    # far outside the map's bounding box
    a = ctx.assoc_bulk(fr, f.pose, THR, 0, pkg.capi.ALGO_GRID).copy()
    b = ctx.assoc_bulk(fr, f.pose, THR, 0, pkg.capi.ALGO_BRUTE).copy()
    assert np.array_equal(a, b) and a[5] == -1 and a[6] == -1
    # a NaN cone in the map (what the reference stores after an az == 0 column) matches nothing
    ctx.map_append(np.array([np.nan]), np.array([np.nan]), np.array([1], dtype=np.int32))
    a2 = ctx.assoc_bulk(fr, f.pose, THR, 0, pkg.capi.ALGO_GRID).copy()
    assert np.array_equal(a2, a)
    # duplicate cones: the LOWEST index wins (first-fit, slam.cpp:575-607)
    ctx.map_clear()
    x = np.concatenate([f.map_x, f.map_x]); y = np.concatenate([f.map_y, f.map_y]); t = np.concatenate([f.map_type, f.map_type])
    ctx.map_append(x, y, t)
    a3 = ctx.assoc_bulk(f.frame, f.pose, THR, 0, pkg.capi.ALGO_GRID).copy()
    assert np.all(a3 < 5000) and (a3 >= 0).sum() > 300


def test_bulk_full_size_c4_grid_equals_brute(ctx, orc, synth, pkg):
    """BASELINE config 4 at full size (1M cones, 100k observations): both device algorithms agree on
    every observation, and a 1,500-observation sample agrees with the oracle."""
    f = synth.cone_field()
    ctx.map_clear()
    ctx.map_append(f.map_x, f.map_y, f.map_type)
    for gate in (0, 1):
        grid = ctx.assoc_bulk(f.frame, f.pose, THR, gate, pkg.capi.ALGO_GRID).copy()
        brute = ctx.assoc_bulk(f.frame, f.pose, THR, gate, pkg.capi.ALGO_BRUTE).copy()
        assert np.array_equal(grid, brute)
        sample = np.asfortranarray(f.frame[:, :1500])
        o = orc.assoc_match_only(sample, f.pose, THR, gate, f.map_x, f.map_y, f.map_type)
        assert np.array_equal(grid[:1500], o["idx"])
    assert (grid >= 0).mean() > 0.85


@pytest.mark.parametrize("algo_name", ["ALGO_GRID_PIPELINED", "ALGO_GRID_BATCHED"])
@pytest.mark.parametrize("gate", [0, 1])
def test_bulk_pipelined_train_equals_single_frames(ctx, orc, synth, pkg, gate, algo_name):
    """SLAM_B200_ALGO_GRID_PIPELINED / _BATCHED: a train of independent frames (different observation sets and
    poses) overlapping on one stream / sharing launches of eight gives, frame by frame, what ALGO_GRID and the
    oracle give (12 frames: one full batch of eight and a ragged one of four; ragged frame sizes in the batch)."""
    import torch
    f = synth.cone_field(n_map=200_000, n_obs=24_000, seed=9)
    ctx.map_clear()
    ctx.map_append(f.map_x, f.map_y, f.map_type)
    ctx.map_build_grid(THR)
    F, n = 12, 2000
    dev = torch.device("cuda", 0)
    poses = np.tile(f.pose, (F, 1))
    poses[:, 0] += np.arange(F) * 0.37          # shifted poses: every frame has its own answer
    poses[:, 2] += np.arange(F) * 0.01
    frames = [np.asfortranarray(f.frame[:, k * n:(k + 1) * n]) for k in range(F)]
    d_in = [torch.from_numpy(np.ascontiguousarray(fr.T)).to(dev) for fr in frames]
    d_out = [torch.full((n,), -7, dtype=torch.int32, device=dev) for _ in range(F)]
    torch.cuda.synchronize()
    launch = pkg.capi.Context.assoc_bulk_frames_dev([ctx] * F, [t.data_ptr() for t in d_in], [n] * F, poses, THR, gate,
                                                    getattr(pkg.capi, algo_name), [t.data_ptr() for t in d_out])
    for _ in range(3):
        assert launch() == F
    ctx.sync()
    matched = 0
    for k in range(F):
        got = d_out[k].cpu().numpy()
        single = ctx.assoc_bulk(frames[k], poses[k], THR, gate, pkg.capi.ALGO_GRID).copy()
        assert np.array_equal(got, single), f"frame {k}"
        if k % 4 == 0:
            o = orc.assoc_match_only(frames[k], poses[k], THR, gate, f.map_x, f.map_y, f.map_type)
            assert np.array_equal(got, o["idx"]), f"frame {k} vs oracle"
        matched += int((got >= 0).sum())
    assert matched > 1000


def test_whole_drives_per_replica_equal_the_oracle(ctx, orc, synth):
    """SURVEY section 7 step 6 / 8(d) 'optional second variant' of config 3: whole Monte-Carlo drives, one replica per
    thread block, frame state carried on the device (slam_b200_drive_replicas).  Seven noisy replicas of a small
    loop (each with its own pose / observation noise, so their maps and the frame that closes the loop differ) + one
    replica with an empty frame, a NaN cone (azimuth 0) and a column beyond the mapping threshold: every frame's
    records, the map and the closing frame must be what the oracle's addConesToMap gives frame by frame."""
    R, F, nmax, cap = 8, 230, 16, 256
    frames = np.zeros((R, F, 4, nmax)); ncols = np.zeros((R, F), dtype=np.int32); poses = np.zeros((R, F, 3))
    for r in range(R):
        trk = synth.ellipse_track(n_pairs=22 + 2 * (r % 3), a=18.0 + 2 * (r % 3), b=9.0 + (r % 3), half_width=1.5)
        d = synth.simulate_drive(trk, F, s_step=1.45 * trk.length / F, seed=100 + r, sigma_xy=0.05 + 0.01 * r, sigma_th=0.005)
        for f, (fr, p) in enumerate(zip(d.frames, d.poses_noisy)):
            fr = np.asarray(fr, dtype=np.float64).reshape(4, -1, order="F")[:, :nmax]
            if r == R - 1:
                if f == 5:
                    fr = fr[:, :0]
                elif f == 9 and fr.shape[1] > 1:
                    fr = fr.copy(); fr[0, 1] = 0.0            # azimuth 0 -> NaN cone
                elif f == 12 and fr.shape[1] > 1:
                    fr = fr.copy(); fr[2, 0] = np.float32(75.0)
            n = fr.shape[1]
            frames[r, f, :, :n] = fr
            ncols[r, f] = n
            poses[r, f] = p
    out = ctx.drive_replicas(frames, ncols, poses, synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD, cap=cap)
    closed_seen = set()
    for r in range(R):
        mx = np.zeros(cap); my = np.zeros(cap); mt = np.zeros(cap, dtype=np.int32)
        M = cci = lc = 0
        closed = -1
        for f in range(F):
            n = int(ncols[r, f])
            sc = out["scalars"][r, f]
            if lc or n == 0:
                assert sc[5] == -1 and sc[2] == M and sc[4] == lc, (r, f)   # recorded as not run, state unchanged
                continue
            o = orc.assoc_map_frame(frames[r, f, :, :n], poses[r, f], synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD,
                                    mx, my, mt, M, cci, lc)
            assert np.array_equal(out["idx"][r, f, :n], o["idx"]) and np.array_equal(out["status"][r, f, :n], o["status"]), (r, f)
            M, cci, lc = o["M"], o["cci"], o["loop_closing"]
            assert sc[2] == M and sc[3] == cci and sc[4] == lc and sc[5] == 0, (r, f)
            if lc and closed < 0:
                closed = f
        assert out["closed_at"][r] == closed and out["map_n"][r] == M, r
        ok = ~np.isnan(mx[:M])
        assert np.array_equal(np.isnan(out["map_x"][r, :M]), ~ok)
        assert np.allclose(out["map_x"][r, :M][ok], mx[:M][ok], rtol=0, atol=1e-9) and np.array_equal(out["map_type"][r, :M], mt[:M])
        closed_seen.add(closed)
    assert len(closed_seen) > 2 and -1 not in closed_seen        # the replicas really differ, and all of them close

"""The C++ Slam mirror (csrc/host/slam.cpp, the reference's class interface over the C ABI) replayed
frame by frame against the restated reference back half in the oracle: configuration C1."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def host(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pkg.build()
    from importlib import import_module
    b = import_module(pkg.__name__ + "._build")
    L = C.CDLL(b.HOSTLIB)
    L.slamhost_create.restype = C.c_void_p
    L.slamhost_create.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
    L.slamhost_destroy.argtypes = [C.c_void_p]
    L.slamhost_perform.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, C.c_float, C.c_double, c_ip, c_ip]
    L.slamhost_last_error.restype = C.c_char_p
    L.slamhost_cone_packet.argtypes = [C.c_void_p, C.c_int, c_ip, C.POINTER(C.c_float), C.POINTER(C.c_float), c_ip]
    for n in ("slamhost_state", "slamhost_chi2_log", "slamhost_draw_cones", "slamhost_draw_poses",
              "slamhost_draw_current_pose", "slamhost_draw_graph", "slamhost_pose_estimate"):
        getattr(L, n).argtypes = None
    return L


class HostSlam:
    def __init__(self, L, thr, map_thr):
        self.L = L
        self.h = C.c_void_p(L.slamhost_create(thr, map_thr, 20, 0))
        assert self.h, L.slamhost_last_error()

    def close(self):
        self.L.slamhost_destroy(self.h)

    def perform(self, frame, pose, yaw=0.0, dt=0.0):
        frame = np.asfortranarray(frame, dtype=np.float64); pose = np.ascontiguousarray(pose, dtype=np.float64)
        n = frame.shape[1]
        idx = np.zeros(max(n, 1), dtype=np.int32); st = np.zeros(max(n, 1), dtype=np.int32)
        rc = self.L.slamhost_perform(self.h, frame.ctypes.data_as(c_dp), n, pose.ctypes.data_as(c_dp), yaw, dt,
                                     idx.ctypes.data_as(c_ip), st.ctypes.data_as(c_ip))
        assert rc != -100, self.L.slamhost_last_error()
        return rc, idx[:n], st[:n]

    def state(self):
        out = np.zeros(8, dtype=np.int32)
        self.L.slamhost_state(self.h, out.ctypes.data_as(c_ip))
        return out

    def chi2(self):
        n = self.state()[6]
        out = np.zeros(max(n, 1))
        self.L.slamhost_chi2_log(self.h, out.ctypes.data_as(c_dp))
        return out[:n]

    def cones(self):
        M = self.state()[7]
        x = np.zeros(max(M, 1)); y = np.zeros(max(M, 1)); t = np.zeros(max(M, 1), dtype=np.int32); i = np.zeros(max(M, 1), dtype=np.int32)
        self.L.slamhost_draw_cones(self.h, x.ctypes.data_as(c_dp), y.ctypes.data_as(c_dp), t.ctypes.data_as(c_ip), i.ctypes.data_as(c_ip))
        return x[:M], y[:M], t[:M], i[:M]

    def poses(self, cap=4096):
        out = np.zeros((cap, 3))
        n = self.L.slamhost_draw_poses(self.h, out.ctypes.data_as(c_dp), cap)
        return out[:n]

    def graph(self, cap=4096, capf=65536):
        cnt = np.zeros(cap, dtype=np.int32); flat = np.zeros(capf, dtype=np.int32)
        n = self.L.slamhost_draw_graph(self.h, cnt.ctypes.data_as(c_ip), cap, flat.ctypes.data_as(c_ip), capf)
        return cnt[:n], flat[:cnt[:n].sum()]

    def pose_estimate(self, vid):
        out = np.zeros(3)
        d = self.L.slamhost_pose_estimate(self.h, int(vid), out.ctypes.data_as(c_dp))
        return d, out

    def current_pose(self):
        out = np.zeros(3)
        self.L.slamhost_draw_current_pose(self.h, out.ctypes.data_as(c_dp))
        return out


def test_c1_replay_through_slam_class(host, orc, synth, c1_drive):
    s = HostSlam(host, synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    o = orc.slam(synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    kinds = []
    for k, (fr, p) in enumerate(zip(c1_drive.frames, c1_drive.poses_noisy)):
        rc, idx, st = s.perform(fr, p)
        rco, idxo, sto = o.perform(fr, p)
        assert rc == rco, k
        assert np.array_equal(idx, idxo) and np.array_equal(st, sto), k     # association bit-exact
        kinds.append(rc)
    assert 1 in kinds and 2 in kinds                                         # closed the loop, then localised
    so = o.state(); sd = s.state()
    assert sd[0] == so["current_cone_index"] and sd[1] == so["pose_id"]
    assert sd[3] == 1 and sd[4] == so["optimize_calls"] and sd[5] == so["last_iterations"] == 10
    assert np.allclose(s.chi2(), o.chi2_log(), rtol=1e-8)
    x, y, t, ids = s.cones(); ox, oy, ot = o.map()
    assert np.array_equal(t, ot) and np.array_equal(ids, np.arange(len(ids)))
    scale = max(1.0, np.abs(ox).max(), np.abs(oy).max())
    assert np.max(np.abs(x - ox)) <= 1e-6 * scale and np.max(np.abs(y - oy)) <= 1e-6 * scale
    # optimised poses (graph vertices), first/last and a few in between
    for vid in [1000, 1001, 1002, 1500, 1973, 1974, 1999]:
        d, e = s.pose_estimate(vid)
        assert d == 3 and np.max(np.abs(e - o.pose(vid))) <= 1e-6 * scale, vid
    assert np.allclose(s.current_pose(), o.send_pose(), rtol=0, atol=1e-6 * scale)
    assert len(s.poses()) == 1000
    cnt, flat = s.graph()
    assert len(cnt) == 1000 and cnt.sum() == so["n_edges"] - 999            # cone edges; odometry edges excluded
    # the packet sendCones() emits (slam.cpp:656-679): 20 cones from m_currentConeIndex on, wrapping
    mi = np.zeros(32, dtype=np.int32); az = np.zeros(32, dtype=np.float32); di = np.zeros(32, dtype=np.float32)
    ty = np.zeros(32, dtype=np.int32)
    n = host.slamhost_cone_packet(s.h, 32, mi.ctypes.data_as(c_ip), az.ctypes.data_as(C.POINTER(C.c_float)),
                                  di.ctypes.data_as(C.POINTER(C.c_float)), ty.ctypes.data_as(c_ip))
    assert n == 20
    M = len(x); cci = int(sd[0]); sp = s.current_pose()
    assert np.array_equal(mi[:n], (cci + np.arange(20)) % M)
    assert np.array_equal(ty[:n], t[mi[:n]])
    want_d = np.hypot(x[mi[:n]] - sp[0], y[mi[:n]] - sp[1]).astype(np.float32)
    assert np.allclose(di[:n], want_d, rtol=1e-6)
    want_az = (np.degrees(np.arctan2(y[mi[:n]] - sp[1], x[mi[:n]] - sp[0])) - sp[2] / 57.295779513082325).astype(np.float32)
    assert np.allclose(az[:n], want_az, rtol=1e-5, atol=1e-4)
    s.close()


def test_c1_replay_equals_the_reference_slam_cpp(host, synth, c1_drive):
    """The drop-in Slam (host mirror over the CUDA back end) against what the reference's REAL src/slam.cpp
    produced on the same drive (tests/golden/c1_replay_reference.npz, made by compiling slam.cpp with the
    g2o facade: tests/golden/make_c1_reference_replay.py): the cones every frame was associated with or
    created (drawGraph rows, 8,381 entries), map size / current-cone index / loop-closure flags per frame
    and the stored poses are identical; optimised map and pose vertices within 1e-6 relative."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_replay_reference.npz"))
    s = HostSlam(host, synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    for k, (fr, p) in enumerate(zip(c1_drive.frames, c1_drive.poses_noisy)):
        s.perform(fr, p)
        st = s.state()
        assert st[7] == g["frame_map_size"][k] and st[0] == g["frame_cci"][k], k
        assert st[2] == g["frame_loop_closing"][k] and st[3] == g["frame_loop_closed"][k] and st[1] == g["frame_pose_id"][k], k
    cnt, flat = s.graph()
    assert np.array_equal(np.concatenate([[0], np.cumsum(cnt)]), g["row_ptr"])
    assert np.array_equal(flat, g["row_ids"])
    assert s.poses().tobytes() == np.ascontiguousarray(g["poses"]).tobytes()
    x, y, t, ids = s.cones()
    assert np.array_equal(t, g["map_type"]) and np.array_equal(ids, g["map_id"])
    scale = max(1.0, np.abs(g["map_x"]).max(), np.abs(g["map_y"]).max())
    assert np.max(np.abs(x - g["map_x"])) <= 1e-6 * scale and np.max(np.abs(y - g["map_y"])) <= 1e-6 * scale
    for k in (0, 1, 2, 500, 973, 974, 999):
        d, e = s.pose_estimate(1000 + k)
        assert d == 3 and np.max(np.abs(e - g["vertices"][k])) <= 1e-6 * scale, k
    s.close()


def test_adversarial_replays_equal_the_reference_slam_cpp(host, synth):
    """The drop-in Slam on the twelve adversarial replays of tests/golden/fuzz_replay_reference.npz (made by
    the reference's real slam.cpp: make_fuzz_reference_replay.py) -- repeated columns, non-integer types,
    out-of-range cones, empty frames, scrambled order, NaN cones, early loop closure + localiser phase:
    association rows and per-frame state identical; map within 1e-9 m where no optimisation ran (device
    sin/cos differ from glibc in the last bits; NaN cones in the same slots), 1e-6 relative after it."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_fuzz_reference_replay import scenarios
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fuzz_replay_reference.npz"))
    for name, frames, poses, thr, map_thr in scenarios(synth):
        s = HostSlam(host, thr, map_thr)
        for k, (fr, p) in enumerate(zip(frames, poses)):
            s.perform(fr, p)
            st = s.state()
            got = (int(st[7]), int(st[0]), int(st[2]), int(st[3]), int(st[1]))
            want = tuple(int(g[name + "/" + key][k]) for key in ("frame_map_size", "frame_cci", "frame_loop_closing", "frame_loop_closed", "frame_pose_id"))
            assert got == want, (name, k, got, want)
        cnt, flat = s.graph()
        assert np.array_equal(np.concatenate([[0], np.cumsum(cnt)]), g[name + "/row_ptr"]), name
        assert np.array_equal(flat, g[name + "/row_ids"]), name
        x, y, t, ids = s.cones()
        assert np.array_equal(t, g[name + "/map_type"]), name
        gx, gy = g[name + "/map_x"], g[name + "/map_y"]
        assert np.array_equal(np.isnan(x), np.isnan(gx)) and np.array_equal(np.isnan(y), np.isnan(gy)), name
        ok = ~np.isnan(gx)
        if g[name + "/frame_loop_closed"].any():
            scale = max(1.0, np.abs(gx[ok]).max(), np.abs(gy[ok]).max())
            assert np.max(np.abs(x[ok] - gx[ok])) <= 1e-6 * scale and np.max(np.abs(y[ok] - gy[ok])) <= 1e-6 * scale, name
        else:
            assert np.max(np.abs(x[ok] - gx[ok])) <= 1e-9 and np.max(np.abs(y[ok] - gy[ok])) <= 1e-9, name
        s.close()


def test_recording_replay_into_slam_equals_direct_calls(host, synth):
    """tests/golden/c1_head.rec (written by the reference's own cluon) replayed through the front half into the
    drop-in Slam gives exactly the state that feeding the same keyframes to performSLAM by hand gives."""
    import os
    import sys
    here = os.path.dirname(__file__)
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_rec_golden as gen
    rec = os.path.join(here, "golden", "c1_head.rec").encode()
    i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    host.slamrec_replay.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                    i32p, c_dp, c_dp, C.POINTER(C.c_float), c_dp, i64p, i64p]
    host.slamrec_replay_into_slam.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_double, C.c_double, C.c_double]
    nc = np.zeros(256, dtype=np.int32); cones = np.zeros(4 * 8192); odo = np.zeros((256, 3)); yaw = np.zeros(256, dtype=np.float32)
    el = np.zeros(256); tu = np.zeros(256, dtype=np.int64); st = np.zeros(9, dtype=np.int64)
    n = host.slamrec_replay(rec, gen.DETECT_CONE_ID, gen.ESTIMATION_ID, 10, 0.5, gen.REF_LAT, gen.REF_LON, 256, 8192,
                            nc.ctypes.data_as(i32p), cones.ctypes.data_as(c_dp), odo.ctypes.data_as(c_dp),
                            yaw.ctypes.data_as(C.POINTER(C.c_float)), el.ctypes.data_as(c_dp), tu.ctypes.data_as(i64p), st.ctypes.data_as(i64p))
    assert n == 40
    a = HostSlam(host, synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    done = host.slamrec_replay_into_slam(rec, a.h, gen.DETECT_CONE_ID, gen.ESTIMATION_ID, 10, 0.5, gen.REF_LAT, gen.REF_LON)
    assert done == 40, host.slamhost_last_error()
    b = HostSlam(host, synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    off = 0
    for k in range(n):
        fr = cones[4 * off:4 * (off + nc[k])].reshape(nc[k], 4).T
        off += nc[k]
        b.perform(fr, odo[k], float(yaw[k]), float(el[k]))
    assert np.array_equal(a.state(), b.state()) and a.state()[7] > 20            # same map size, indices, flags
    for u, v in zip(a.cones(), b.cones()):
        assert np.array_equal(u, v)
    ca, fa = a.graph(); cb, fb = b.graph()
    assert np.array_equal(ca, cb) and np.array_equal(fa, fb) and len(ca) == 40
    assert a.poses().tobytes() == b.poses().tobytes()
    a.close(); b.close()


def test_heading_correction_replays_equal_the_reference_slam_cpp(host, synth):
    """performSLAM's heading correction (slam.cpp:309-318) through the drop-in Slam (setYawRate) against the
    reference's real slam.cpp: tests/golden/fuzz_yaw_replay_reference.npz."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_fuzz_reference_replay import scenarios_yaw
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fuzz_yaw_replay_reference.npz"))
    for name, frames, poses, thr, map_thr, yaw in scenarios_yaw(synth):
        s = HostSlam(host, thr, map_thr)
        for k, (fr, p) in enumerate(zip(frames, poses)):
            s.perform(fr, p, float(yaw[k][0]), abs(float(yaw[k][1])) / 1000000)
            st = s.state()
            got = (int(st[7]), int(st[0]), int(st[2]), int(st[3]), int(st[1]))
            want = tuple(int(g[name + "/" + key][k]) for key in ("frame_map_size", "frame_cci", "frame_loop_closing", "frame_loop_closed", "frame_pose_id"))
            assert got == want, (name, k, got, want)
        cnt, flat = s.graph()
        assert np.array_equal(np.concatenate([[0], np.cumsum(cnt)]), g[name + "/row_ptr"]) and np.array_equal(flat, g[name + "/row_ids"]), name
        assert s.poses().tobytes() == np.ascontiguousarray(g[name + "/poses"]).tobytes(), name
        s.close()


def test_burst_of_optimise_calls_and_gates(host, orc, synth, c1_drive):
    """Closing column first in its frame -> one optimise per remaining column (slam.cpp:625-633);
    a pose outside +-200 m is rejected (300-303); the yaw-rate heading correction (315-317)."""
    frames = [f.copy(order="F") for f in c1_drive.frames]
    probe = orc.slam(synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    for k, (fr, p) in enumerate(zip(frames, c1_drive.poses_noisy)):
        if probe.perform(fr, p)[0] == 1:
            break
    frames[k] = np.asfortranarray(frames[k][:, ::-1])
    s = HostSlam(host, synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    o = orc.slam(synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    assert s.perform(frames[0], np.array([250.0, 0.0, 0.0]))[0] == -1
    assert o.perform(frames[0], np.array([250.0, 0.0, 0.0]))[0] == -1
    for q in range(k + 1):
        yaw, dt = (0.05, 0.2) if q % 5 == 0 else (0.0, 0.0)
        rc, idx, st = s.perform(frames[q], c1_drive.poses_noisy[q], yaw, dt)
        rco, idxo, sto = o.perform(frames[q], c1_drive.poses_noisy[q], np.float32(yaw), dt)
        assert rc == rco and np.array_equal(idx, idxo) and np.array_equal(st, sto), q
    n_cols = frames[k].shape[1]
    assert s.state()[4] == o.state()["optimize_calls"] == n_cols            # closing column was column 0
    assert np.allclose(s.chi2(), o.chi2_log(), rtol=1e-7)
    x, y, _, _ = s.cones(); ox, oy, _ = o.map()
    assert np.allclose(x, ox, rtol=0, atol=1e-4) and np.allclose(y, oy, rtol=0, atol=1e-4)
    s.close()


def test_constructor_and_cone_accessors(host):
    assert host.slamhost_create_missing_key_throws() == 1                  # std::stoi throws (slam.cpp:739)
    az = C.c_float(); dist = C.c_float()
    pose = np.array([1.0, 2.0, 30.0])
    host.slamhost_cone_bearing(C.c_double(4.0), C.c_double(6.0), pose.ctypes.data_as(c_dp), C.byref(az), C.byref(dist))
    assert dist.value == pytest.approx(5.0)
    want = np.degrees(np.arctan2(4.0, 3.0)) - 30.0 / 57.295779513082325      # src/cone.cpp:37-39
    assert az.value == pytest.approx(np.float32(want), rel=1e-6)



@pytest.mark.parametrize("window", [3, 10])
def test_localizer_repair_equals_the_oracle(host, window):
    """localiser frames with observation edges + sliding-window optimise through the
    drop-in Slam (setLocalizerRepair) against the restated semantics in the oracle, two short laps.  Runs in its own
    process (tests/gpu_case_localizer_repair.py) so that a device fault on this not-yet-measured path cannot poison
    the CUDA context of the tests after it."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "gpu_case_localizer_repair.py"), str(window)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_gate_replays_equal_the_reference_slam_cpp(host, synth):
    """The drop-in Slam on tests/golden/fuzz_gate_replay_reference.npz (the reference's real slam.cpp): loops that
    partly lie beyond the 200 m gate of performSLAM (frames dropped before a pose is added), odd cone types, zero
    ranges, azimuths beyond the half circle, zenith != 0, single-column frames."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_fuzz_reference_replay import scenarios_gate
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fuzz_gate_replay_reference.npz"))
    for name, frames, poses, thr, map_thr in scenarios_gate(synth):
        s = HostSlam(host, thr, map_thr)
        present = g[name + "/row_present"]
        for k, (fr, p) in enumerate(zip(frames, poses)):
            rc, _, _ = s.perform(fr, p)
            st = s.state()
            got = (int(st[7]), int(st[0]), int(st[2]), int(st[3]), int(st[1]))
            want = tuple(int(g[name + "/" + key][k]) for key in ("frame_map_size", "frame_cci", "frame_loop_closing", "frame_loop_closed", "frame_pose_id"))
            assert got == want, (name, k, got, want)
            assert (rc == -1) == (not present[k]), (name, k, rc)
        cnt, flat = s.graph()
        assert np.array_equal(np.concatenate([[0], np.cumsum(cnt)]), g[name + "/row_ptr"][np.concatenate([[True], present])]), name
        assert np.array_equal(flat, g[name + "/row_ids"]), name
        x, y, t, ids = s.cones()
        assert np.array_equal(t, g[name + "/map_type"]), name
        gx, gy = g[name + "/map_x"], g[name + "/map_y"]
        tol = 1e-6 * max(1.0, np.abs(gx).max(), np.abs(gy).max()) if g[name + "/frame_loop_closed"].any() else 1e-9
        assert np.max(np.abs(x - gx)) <= tol and np.max(np.abs(y - gy)) <= tol, name
        s.close()

"""ctypes binding of the CPU oracle (oracle/slam_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

`load(kind)`: kind "port" -> oracle/liboracle.so (own LDL^T); "reference" ->
oracle/_ref/liboracle_eigen.so (the reference's vendored Eigen 3.3.4 SimplicialLDLT+AMD);
"best" -> reference if present else port.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "liboracle_eigen.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(quiet=True):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


class Oracle:
    def __init__(self, path, kind):
        self.kind = kind
        self.path = path
        L = self.lib = C.CDLL(path)
        L.orc_pi_ref.restype = C.c_double
        L.orc_normalize_theta.restype = C.c_double
        L.orc_normalize_theta.argtypes = [C.c_double]
        L.orc_graph_create.restype = C.c_void_p
        L.orc_graph_chi2.restype = C.c_double
        L.orc_graph_chi2.argtypes = [C.c_void_p]
        L.orc_graph_build_system.restype = C.c_long
        L.orc_slam_create.restype = C.c_void_p
        L.orc_slam_create.argtypes = [C.c_double, C.c_double]
        L.orc_slam_graph.restype = C.c_void_p
        L.orc_slam_graph.argtypes = [C.c_void_p]
        for name in ("orc_graph_destroy", "orc_slam_destroy"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_transform_cone_to_cog.argtypes = [C.c_double, C.c_double, c_dp]
        L.orc_spherical2cartesian.argtypes = [C.c_double, C.c_double, C.c_double, c_dp]
        L.orc_graph_add_pose.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_graph_add_landmark.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.orc_slam_perform.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, C.c_double, C.c_double, c_ip, c_ip]
        L.orc_slam_set_localizer_repair.argtypes = [C.c_void_p, C.c_int, C.c_int]

    # ---- conversions -----------------------------------------------------------------------
    def transform_cone_to_cog(self, angle, distance):
        out = np.zeros(2)
        self.lib.orc_transform_cone_to_cog(angle, distance, _dp(out))
        return out

    def spherical2cartesian(self, az, zen, d):
        out = np.zeros(3)
        self.lib.orc_spherical2cartesian(az, zen, d, _dp(out))
        return out

    def cone_to_global(self, pose, obs4):
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        obs4 = np.ascontiguousarray(obs4, dtype=np.float64)
        out = np.zeros(3)
        self.lib.orc_cone_to_global(_dp(pose), _dp(obs4), _dp(out))
        return out

    def odometry_measurement(self, prev, cur):
        prev = np.ascontiguousarray(prev, dtype=np.float64)
        cur = np.ascontiguousarray(cur, dtype=np.float64)
        out = np.zeros(3)
        self.lib.orc_odometry_measurement(_dp(prev), _dp(cur), _dp(out))
        return out

    # ---- association -------------------------------------------------------------------------
    def assoc_map_frame(self, frame, pose, thr, map_thr, map_x, map_y, map_type, M, cci, loop_closing):
        """Arrays map_* must have capacity >= M + N + 1 and are updated in place.
        Returns dict(idx, status, z, g, first, lc_obs, M, cci, loop_closing)."""
        frame = np.asfortranarray(frame, dtype=np.float64)
        N = frame.shape[1]
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        idx = np.zeros(max(N, 1), dtype=np.int32); st = np.zeros(max(N, 1), dtype=np.int32)
        z = np.zeros((max(N, 1), 2)); g = np.zeros((max(N, 1), 3))
        m = C.c_int(M); c = C.c_uint(cci); lc = C.c_int(loop_closing)
        first = C.c_int(0); lcobs = C.c_int(-1)
        rc = self.lib.orc_assoc_map_frame(_dp(frame), N, _dp(pose), C.c_double(thr), C.c_double(map_thr),
                                          _dp(map_x), _dp(map_y), _ip(map_type), C.byref(m), len(map_x),
                                          C.byref(c), C.byref(lc), _ip(idx), _ip(st), _dp(z), _dp(g),
                                          C.byref(first), C.byref(lcobs))
        assert rc == 0
        return dict(idx=idx[:N], status=st[:N], z=z[:N], g=g[:N], first=first.value, lc_obs=lcobs.value,
                    M=m.value, cci=c.value, loop_closing=lc.value)

    def assoc_localize_frame(self, frame, pose, thr, map_x, map_y, map_type, cci):
        frame = np.asfortranarray(frame, dtype=np.float64)
        N = frame.shape[1]
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        idx = np.zeros(max(N, 1), dtype=np.int32); g = np.zeros((max(N, 1), 3))
        c = C.c_uint(cci); reobs = C.c_int(0); send = C.c_int(0)
        self.lib.orc_assoc_localize_frame(_dp(frame), N, _dp(pose), C.c_double(thr), _dp(map_x), _dp(map_y),
                                          _ip(map_type), len(map_x), C.byref(c), _ip(idx), _dp(g),
                                          C.byref(reobs), C.byref(send))
        return dict(idx=idx[:N], g=g[:N], cci=c.value, n_reobserved=reobs.value, send_cone_data=send.value)

    def assoc_match_only(self, frame, pose, thr, mode, map_x, map_y, map_type, want_g=False):
        frame = np.asfortranarray(frame, dtype=np.float64)
        N = frame.shape[1]
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        idx = np.zeros(max(N, 1), dtype=np.int32)
        g = np.zeros((max(N, 1), 3)) if want_g else None
        mm = C.c_double(0)
        self.lib.orc_assoc_match_only(_dp(frame), N, _dp(pose), C.c_double(thr), mode, _dp(map_x), _dp(map_y),
                                      _ip(map_type), len(map_x), _ip(idx), _dp(g), C.byref(mm))
        return dict(idx=idx[:N], g=None if g is None else g[:N], min_margin=mm.value)

    # ---- graph -------------------------------------------------------------------------------
    def graph(self):
        return OracleGraph(self)

    def graph_from_soa(self, g):
        G = OracleGraph(self)
        G.load_soa(g)
        return G

    def slam(self, same_cone_threshold, cone_mapping_threshold):
        return OracleSlam(self, same_cone_threshold, cone_mapping_threshold)


class OracleGraph:
    def __init__(self, orc, handle=None):
        self.o = orc
        self.L = orc.lib
        self.owned = handle is None
        self.h = C.c_void_p(self.L.orc_graph_create()) if handle is None else C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "owned", False) and self.h:
            self.L.orc_graph_destroy(self.h)
            self.h = None

    def add_pose(self, vid, x, y, th):
        return self.L.orc_graph_add_pose(self.h, int(vid), float(x), float(y), float(th))

    def add_landmark(self, vid, x, y):
        return self.L.orc_graph_add_landmark(self.h, int(vid), float(x), float(y))

    def add_edge_se2(self, a, b, z, info):
        z = np.ascontiguousarray(z, dtype=np.float64); info = np.ascontiguousarray(info, dtype=np.float64)
        return self.L.orc_graph_add_edge_se2(self.h, int(a), int(b), _dp(z), _dp(info))

    def add_odometry(self, a, b, pose, info):
        pose = np.ascontiguousarray(pose, dtype=np.float64); info = np.ascontiguousarray(info, dtype=np.float64)
        return self.L.orc_graph_add_odometry(self.h, int(a), int(b), _dp(pose), _dp(info))

    def add_edge_se2_xy(self, p, l, z, info):
        z = np.ascontiguousarray(z, dtype=np.float64); info = np.ascontiguousarray(info, dtype=np.float64)
        return self.L.orc_graph_add_edge_se2_xy(self.h, int(p), int(l), _dp(z), _dp(info))

    def set_fixed(self, vid, flag=True):
        return self.L.orc_graph_set_fixed(self.h, int(vid), int(bool(flag)))

    def load_soa(self, g):
        """Vertices: landmarks then poses (the Hessian order is by id anyway, slam.hpp:118); edges
        interleaved the way performSLAM inserts them (pose k's odometry edge, then its cone edges)."""
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        a = [i32(g.pose_ids), f64(g.pose_est), i32(g.lm_ids), f64(g.lm_est), i32(g.eo_from), i32(g.eo_to), f64(g.eo_z),
             f64(g.eo_info), i32(g.el_pose), i32(g.el_lm), f64(g.el_z), f64(g.el_info), i32(g.fixed_ids)]
        rc = self.L.orc_graph_load(self.h, len(a[0]), _ip(a[0]), _dp(a[1]), len(a[2]), _ip(a[2]), _dp(a[3]),
                                   len(a[4]), _ip(a[4]), _ip(a[5]), _dp(a[6]), _dp(a[7]),
                                   len(a[8]), _ip(a[8]), _ip(a[9]), _dp(a[10]), _dp(a[11]), len(a[12]), _ip(a[12]))
        assert rc == 0

    def optimize(self, iters=10):
        chi2 = np.zeros(max(iters, 1))
        n = self.L.orc_graph_optimize(self.h, int(iters), _dp(chi2))
        return n, chi2[:max(n, 0)]

    def chi2(self):
        return self.L.orc_graph_chi2(self.h)

    def stats(self):
        out = np.zeros(12)
        self.L.orc_graph_stats(self.h, _dp(out))
        keys = ["t_init", "t_struct", "t_linearize", "t_analyze", "t_factor", "t_solve", "t_update", "t_total",
                "nfree", "nnzH", "nnzL", "flopsL"]
        return dict(zip(keys, out.tolist()))

    def get_vertex(self, vid):
        out = np.zeros(3)
        d = self.L.orc_graph_get_vertex(self.h, int(vid), _dp(out))
        if d < 0:
            raise KeyError(vid)
        return out[:d]

    def get_all(self):
        n = self.L.orc_graph_num_vertices(self.h)
        ids = np.zeros(n, dtype=np.int32); out = np.zeros((n, 3))
        self.L.orc_graph_get_all(self.h, _ip(ids), _dp(out), n)
        return ids, out

    def estimates(self, g):
        """(pose_est (P,3), lm_est (L,2)) in the SoA's order."""
        ids, est = self.get_all()
        m = {int(i): k for k, i in enumerate(ids)}
        pe = np.stack([est[m[int(v)]] for v in g.pose_ids])
        le = np.stack([est[m[int(v)], :2] for v in g.lm_ids])
        return pe, le

    def build_system(self):
        """(n, hidx per vertex in insertion order, scipy-ready upper CSC (Ap, Ai, Ax), b, chi2)."""
        nv = self.L.orc_graph_num_vertices(self.h)
        n = C.c_int(0)
        hidx = np.zeros(nv, dtype=np.int32)
        nnz = self.L.orc_graph_build_system(self.h, C.byref(n), _ip(hidx), None, None, None, None, None)
        Ap = np.zeros(n.value + 1, dtype=np.int32); Ai = np.zeros(nnz, dtype=np.int32)
        Ax = np.zeros(nnz); b = np.zeros(n.value); chi2 = C.c_double(0)
        self.L.orc_graph_build_system(self.h, C.byref(n), _ip(hidx), _ip(Ap), _ip(Ai), _dp(Ax), _dp(b), C.byref(chi2))
        return dict(n=n.value, hidx=hidx, Ap=Ap, Ai=Ai, Ax=Ax, b=b, chi2=chi2.value)

    def edge_linearization(self, e):
        err = np.zeros(3); Ji = np.zeros(9); Jj = np.zeros(9)
        D = self.L.orc_graph_edge_linearization(self.h, int(e), _dp(err), _dp(Ji), _dp(Jj))
        return D, err, Ji.reshape(3, 3), Jj.reshape(3, 3)


class OracleSlam:
    """The reference's Slam back half (performSLAM and below) on the oracle."""

    def __init__(self, orc, thr, map_thr):
        self.o = orc
        self.L = orc.lib
        self.h = C.c_void_p(self.L.orc_slam_create(thr, map_thr))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_slam_destroy(self.h)
            self.h = None

    def set_localizer_repair(self, on=True, window=10):
        """SURVEY 8(f) rank 3, opt-in: observation edges + sliding-window optimise in localiser frames."""
        self.L.orc_slam_set_localizer_repair(self.h, int(bool(on)), int(window))

    def perform(self, frame, pose, yaw_rate=0.0, time_elapsed=0.0):
        frame = np.asfortranarray(frame, dtype=np.float64)
        N = frame.shape[1]
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        idx = np.zeros(max(N, 1), dtype=np.int32); st = np.zeros(max(N, 1), dtype=np.int32)
        rc = self.L.orc_slam_perform(self.h, _dp(frame), N, _dp(pose), yaw_rate, time_elapsed, _ip(idx), _ip(st))
        return rc, idx[:N], st[:N]

    def state(self):
        out = np.zeros(8, dtype=np.int32)
        self.L.orc_slam_state(self.h, _ip(out))
        keys = ["current_cone_index", "pose_id", "loop_closing", "loop_closing_complete", "optimize_calls",
                "last_iterations", "n_chi2", "n_edges"]
        return dict(zip(keys, out.tolist()))

    def connectivity_row(self, k):
        out = np.zeros(4096, dtype=np.int32)
        n = self.L.orc_slam_connectivity_row(self.h, int(k), _ip(out), len(out))
        return None if n < 0 else out[:n].copy()

    def poses(self):
        n = self.L.orc_slam_num_poses(self.h)
        out = np.zeros((max(n, 1), 3))
        self.L.orc_slam_get_poses(self.h, _dp(out))
        return out[:n]

    def chi2_log(self):
        n = self.state()["n_chi2"]
        out = np.zeros(max(n, 1))
        self.L.orc_slam_chi2_log(self.h, _dp(out))
        return out[:n]

    def map(self):
        M = self.L.orc_slam_map_size(self.h)
        x = np.zeros(max(M, 1)); y = np.zeros(max(M, 1)); t = np.zeros(max(M, 1), dtype=np.int32)
        self.L.orc_slam_get_map(self.h, _dp(x), _dp(y), _ip(t))
        return x[:M], y[:M], t[:M]

    def pose(self, vid):
        out = np.zeros(3)
        if self.L.orc_slam_get_pose(self.h, int(vid), _dp(out)) < 0:
            raise KeyError(vid)
        return out

    def send_pose(self):
        out = np.zeros(3)
        self.L.orc_slam_send_pose(self.h, _dp(out))
        return out

    def graph(self):
        return OracleGraph(self.o, handle=self.L.orc_slam_graph(self.h))


_cache = {}


def load(kind="best") -> Oracle:
    if kind == "best":
        kind = "reference" if os.path.exists(_REF) else "port"
    if kind not in _cache:
        path = _REF if kind == "reference" else _PORT
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        _cache[kind] = Oracle(path, kind)
    return _cache[kind]


def have_reference():
    return os.path.exists(_REF)


def reference_replay_timing(frames, poses, thr, map_thr):
    """Wall time the reference's REAL src/slam.cpp (oracle/_ref/ref_slam_replay, built by build_ref_slam.sh)
    spends inside performSLAM on a drive, by frame kind.  None when the binary is not there."""
    import struct
    import tempfile
    exe = os.path.join(_HERE, "_ref", "ref_slam_replay")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "frames.bin"), os.path.join(tmp, "out.txt")
        with open(fin, "wb") as f:
            f.write(struct.pack("<idd", len(frames), thr, map_thr))
            for fr, p in zip(frames, poses):
                fr = np.asfortranarray(fr, dtype=np.float64)
                f.write(np.asarray(p, dtype=np.float64).tobytes())
                f.write(struct.pack("<i", fr.shape[1]))
                f.write(fr.tobytes(order="F"))
        try:
            subprocess.run([exe, fin, fout], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
        except Exception:
            return None
        for ln in open(fout):
            t = ln.split()
            if t and t[0] == "T":
                nm, sm, nc, sc, nl, sl = int(t[1]), float(t[2]), int(t[3]), float(t[4]), int(t[5]), float(t[6])
                return {"mapping_frames": nm, "us_per_mapping_frame": sm / max(nm, 1) * 1e6,
                        "loop_closing_frames": nc, "ms_per_loop_closing_frame": sc / max(nc, 1) * 1e3,
                        "localiser_frames": nl, "us_per_localiser_frame": sl / max(nl, 1) * 1e6,
                        "whole_drive_ms": (sm + sc + sl) * 1e3,
                        "what": "the reference's real src/slam.cpp (g2o facade over the restated Gauss-Newton), 1 core, "
                                "time inside performSLAM, its per-observation prints sent to /dev/null"}
    return None

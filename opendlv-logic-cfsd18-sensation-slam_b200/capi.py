"""ctypes binding of the C ABI declared in include/slam_b200.h (libslam_b200.so).

This is the call path tests and bench use: Python -> C ABI -> CUDA.  There is no fallback: if the
library is missing it is built with nvcc; if that fails, or no CUDA device is present, the calls
raise.  Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from . import _build

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)

E_CUDA, E_ARG, E_STATE = -100, -101, -102
ASSOC_MATCHED, ASSOC_NEW, ASSOC_NONE, ASSOC_SKIPPED = 0, 1, 2, 3
GATE_MAPPING, GATE_LOCALIZER = 0, 1
ALGO_BRUTE, ALGO_GRID, ALGO_GRID_PIPELINED, ALGO_GRID_BATCHED = 0, 1, 2, 3

_lib = None


class SlamB200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_build.LIB):
            _build.build()
        L = C.CDLL(_build.LIB)
        L.slam_b200_last_error.restype = C.c_char_p
        L.slam_b200_last_error.argtypes = [C.c_void_p]
        L.slam_b200_map_update_from_graph.argtypes = [C.c_void_p]
        L.slam_b200_map_mirror.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double)),
                                           C.POINTER(C.POINTER(C.c_int32))]
        L.slam_b200_warmup.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.slam_b200_drive_replicas.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp, c_ip, c_dp, C.c_double, C.c_double,
                                               C.c_int, c_ip, c_dp, c_dp, c_ip, c_ip, c_ip, C.POINTER(C.c_double)]
        L.slam_b200_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.slam_b200_launch_count.restype = C.c_long
        L.slam_b200_graph_export_system.restype = C.c_long
        L.slam_b200_graph_export_symbolic.restype = C.c_long
        L.slam_b200_graph_system_dev.restype = C.c_long
        L.slam_b200_symbolic_create.restype = C.c_void_p
        L.slam_b200_symbolic_export.restype = C.c_long
        L.slam_b200_symbolic_export.argtypes = [C.c_void_p, C.c_int, c_ip, C.c_long]
        L.slam_b200_symbolic_tileplan.restype = C.c_long
        L.slam_b200_symbolic_tileplan.argtypes = [C.c_void_p, C.c_int, c_ip, C.c_long]
        L.slam_b200_symbolic_stat.restype = C.c_double
        L.slam_b200_symbolic_stat.argtypes = [C.c_void_p, C.c_int]
        L.slam_b200_symbolic_destroy.argtypes = [C.c_void_p]
        for name in ("slam_b200_destroy", "slam_b200_sync", "slam_b200_map_clear", "slam_b200_map_size",
                     "slam_b200_graph_clear", "slam_b200_graph_prepare", "slam_b200_graph_prepare_assembly_only", "slam_b200_graph_num_poses",
                     "slam_b200_graph_num_landmarks", "slam_b200_graph_num_edges", "slam_b200_launch_count",
                     "slam_b200_graph_reset_device", "slam_b200_graph_solve_async", "slam_b200_graph_snapshot",
                     "slam_b200_graph_restore_async"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.slam_b200_graph_add_pose.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.slam_b200_graph_add_landmark.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.slam_b200_graph_add_edge_se2.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp]
        L.slam_b200_graph_add_odometry.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp]
        L.slam_b200_graph_add_edge_se2_xy.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp]
        L.slam_b200_graph_set_fixed.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.slam_b200_graph_optimize.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.slam_b200_graph_iterate_async.argtypes = [C.c_void_p, C.c_int]
        L.slam_b200_batch_iterate_async.argtypes = [C.c_void_p, C.c_int]
        L.slam_b200_graph_finish.argtypes = [C.c_void_p, c_dp, C.c_int]
        L.slam_b200_graph_assemble_async.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.slam_b200_graph_assemble_exchange_async.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.slam_b200_graph_shard_landmarks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.slam_b200_xchg_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ubyte)]
        L.slam_b200_xchg_connect.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte), c_ip]
        L.slam_b200_xchg_error.argtypes = [C.c_void_p]
        L.slam_b200_xchg_set_timeout_ms.argtypes = [C.c_void_p, C.c_double]
        L.slam_b200_map_build_grid.argtypes = [C.c_void_p, C.c_double]
        L.slam_b200_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.slam_b200_profile_read.argtypes = [C.c_void_p, c_dp]
        L.slam_b200_fp64_peak.argtypes = [C.c_void_p, c_dp]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def _f64(a, order="C"):
    return np.require(a, dtype=np.float64, requirements=["C" if order == "C" else "F", "A"])


def _i32(a):
    return np.require(a, dtype=np.int32, requirements=["C", "A"])


def exported_symbols():
    """Every function include/slam_b200.h declares (parsed from the header)."""
    import re
    hdr = os.path.join(os.path.dirname(_build.HERE), "include", "slam_b200.h")
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(slam_b200_[a-z0-9_]+)\s*\(", txt)))


class Context:
    """One slam_b200_ctx.  `stream`: an integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)
    or None for a context-owned stream."""

    def __init__(self, device=0, stream=None):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.slam_b200_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise SlamB200Error(f"slam_b200_create failed ({rc}): no usable CUDA device {device}; "
                                "this library has no CPU fallback")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.slam_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what=""):
        if rc < 0:
            msg = self.L.slam_b200_last_error(self.h)
            raise SlamB200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
        return rc

    def sync(self):
        self._ck(self.L.slam_b200_sync(self.h), "sync")

    def launch_count(self):
        return int(self.L.slam_b200_launch_count(self.h))

    # ---- conversion ----------------------------------------------------------------------------
    def cones_to_global(self, frame, pose):
        frame = _f64(frame, "F"); pose = _f64(pose)
        n = frame.shape[1]
        g = np.zeros((max(n, 1), 3)); l = np.zeros((max(n, 1), 3))
        self._ck(self.L.slam_b200_cones_to_global(self.h, _dp(frame), n, _dp(pose), _dp(g), _dp(l)), "cones_to_global")
        return g[:n], l[:n]

    # ---- map -------------------------------------------------------------------------------------
    def map_clear(self):
        self._ck(self.L.slam_b200_map_clear(self.h))

    def map_append(self, x, y, t):
        x = _f64(x); y = _f64(y); t = _i32(t)
        return self._ck(self.L.slam_b200_map_append(self.h, _dp(x), _dp(y), _ip(t), len(x)), "map_append")

    def map_size(self):
        return self._ck(self.L.slam_b200_map_size(self.h))

    def map_read(self):
        m = self.map_size()
        x = np.zeros(max(m, 1)); y = np.zeros(max(m, 1)); t = np.zeros(max(m, 1), dtype=np.int32)
        self._ck(self.L.slam_b200_map_read(self.h, 0, m, _dp(x), _dp(y), _ip(t)), "map_read")
        return x[:m], y[:m], t[:m]

    def map_write_xy(self, first, x, y):
        x = _f64(x); y = _f64(y)
        self._ck(self.L.slam_b200_map_write_xy(self.h, int(first), len(x), _dp(x), _dp(y)), "map_write_xy")

    def graph_num_poses(self):
        return self._ck(self.L.slam_b200_graph_num_poses(self.h), "graph_num_poses")

    def graph_num_landmarks(self):
        return self._ck(self.L.slam_b200_graph_num_landmarks(self.h), "graph_num_landmarks")

    def map_update_from_graph(self):
        """Slam::updateMap on the device: map cone j <- landmark vertex j; returns the cones updated."""
        return self._ck(self.L.slam_b200_map_update_from_graph(self.h), "map_update_from_graph")

    def map_mirror(self):
        """Copies of the pinned host mirror of the cone map: x, y, type."""
        px = C.POINTER(C.c_double)(); py = C.POINTER(C.c_double)(); pt = C.POINTER(C.c_int32)()
        n = self._ck(self.L.slam_b200_map_mirror(self.h, C.byref(px), C.byref(py), C.byref(pt)), "map_mirror")
        if n == 0:
            return np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int32)
        return (np.ctypeslib.as_array(px, shape=(n,)).copy(), np.ctypeslib.as_array(py, shape=(n,)).copy(),
                np.ctypeslib.as_array(pt, shape=(n,)).copy())

    def warmup(self, poses_hint=1024, landmarks_hint=340):
        self._ck(self.L.slam_b200_warmup(self.h, int(poses_hint), int(landmarks_hint)), "warmup")

    def drive_replicas(self, frames, ncols, poses, thr, map_thr, cap=512):
        """Whole drives per replica (slam_b200_drive_replicas).  frames: (R, F, 4, nmax) with the columns of every
        frame in [..., :n]; ncols: (R, F); poses: (R, F, 3).  Returns a dict: scalars (R, F, 8), idx / status (R, F, nmax),
        map_x / map_y / map_type (R, cap), map_n (R,), closed_at (R,), kernel_ms."""
        frames = np.asarray(frames, dtype=np.float64)
        R, F, four, nmax = frames.shape
        assert four == 4
        packed = np.ascontiguousarray(np.transpose(frames, (0, 1, 3, 2)))     # (R, F, nmax, 4): column-major 4 x nmax
        ncols = np.ascontiguousarray(ncols, dtype=np.int32); poses = np.ascontiguousarray(poses, dtype=np.float64)
        rec = np.zeros((R, F, 2 * nmax + 8), dtype=np.int32)
        mx = np.zeros((R, cap)); my = np.zeros((R, cap)); mt = np.zeros((R, cap), dtype=np.int32)
        mn = np.zeros(R, dtype=np.int32); closed = np.zeros(R, dtype=np.int32)
        ms = C.c_double(0)
        self._ck(self.L.slam_b200_drive_replicas(self.h, R, F, nmax, _dp(packed), _ip(ncols), _dp(poses), float(thr),
                                                 float(map_thr), int(cap), _ip(rec), _dp(mx), _dp(my), _ip(mt), _ip(mn),
                                                 _ip(closed), C.byref(ms)), "drive_replicas")
        idx = np.full((R, F, nmax), -1, dtype=np.int32); status = np.zeros((R, F, nmax), dtype=np.int32)
        for r in range(R):
            for f in range(F):
                n = min(int(ncols[r, f]), nmax)
                idx[r, f, :n] = rec[r, f, 8:8 + n]
                status[r, f, :n] = rec[r, f, 8 + n:8 + 2 * n]
        return dict(scalars=rec[:, :, :8].copy(), idx=idx, status=status, map_x=mx, map_y=my, map_type=mt, map_n=mn,
                    closed_at=closed, kernel_ms=ms.value)

    def map_build_grid(self, cell):
        return self._ck(self.L.slam_b200_map_build_grid(self.h, float(cell)), "map_build_grid")

    # ---- association -----------------------------------------------------------------------------
    def assoc_map_frame(self, frame, pose, thr, map_thr, cci, loop_closing):
        frame = _f64(frame, "F"); pose = _f64(pose)
        n = frame.shape[1]
        idx = np.zeros(max(n, 1), dtype=np.int32); st = np.zeros(max(n, 1), dtype=np.int32)
        z = np.zeros((max(n, 1), 2)); g = np.zeros((max(n, 1), 3))
        c = C.c_uint32(cci); lc = C.c_int32(loop_closing); first = C.c_int32(0); lcobs = C.c_int32(-1)
        m = self._ck(self.L.slam_b200_assoc_map_frame(self.h, _dp(frame), n, _dp(pose), C.c_double(thr),
                                                      C.c_double(map_thr), C.byref(c), C.byref(lc), _ip(idx),
                                                      _ip(st), _dp(z), _dp(g), C.byref(first), C.byref(lcobs)),
                     "assoc_map_frame")
        return dict(idx=idx[:n], status=st[:n], z=z[:n], g=g[:n], first=first.value, lc_obs=lcobs.value, M=m,
                    cci=c.value, loop_closing=lc.value)

    def assoc_localize_frame(self, frame, pose, thr, cci):
        frame = _f64(frame, "F"); pose = _f64(pose)
        n = frame.shape[1]
        idx = np.zeros(max(n, 1), dtype=np.int32); g = np.zeros((max(n, 1), 3))
        c = C.c_uint32(cci); reobs = C.c_int32(0); send = C.c_int32(0)
        self._ck(self.L.slam_b200_assoc_localize_frame(self.h, _dp(frame), n, _dp(pose), C.c_double(thr), C.byref(c),
                                                       _ip(idx), _dp(g), C.byref(reobs), C.byref(send)),
                 "assoc_localize_frame")
        return dict(idx=idx[:n], g=g[:n], cci=c.value, n_reobserved=reobs.value, send_cone_data=send.value)

    def assoc_bulk(self, frame, pose, thr, gate=GATE_MAPPING, algo=ALGO_GRID, out=None):
        """Host buffers in, host idx out (H2D + kernel + D2H inside the call)."""
        frame = _f64(frame, "F"); pose = _f64(pose)
        n = frame.shape[1]
        idx = out if out is not None else np.zeros(max(n, 1), dtype=np.int32)
        self._ck(self.L.slam_b200_assoc_bulk(self.h, _dp(frame), n, _dp(pose), C.c_double(thr), gate, algo, _ip(idx)),
                 "assoc_bulk")
        return idx[:n]

    def assoc_bulk_dev(self, cones_dev_ptr, n, pose, thr, gate, algo, idx_dev_ptr):
        pose = _f64(pose)
        self._ck(self.L.slam_b200_assoc_bulk_dev(self.h, C.c_void_p(cones_dev_ptr), int(n), _dp(pose), C.c_double(thr),
                                                 gate, algo, C.c_void_p(idx_dev_ptr)), "assoc_bulk_dev")

    @staticmethod
    def assoc_bulk_frames_dev(ctxs, cones_dev_ptrs, ns, poses, thr, gate, algo, idx_dev_ptrs):
        """A train of independent frames, one launch each (see include/slam_b200.h); returns a callable
        that relaunches the same train without rebuilding the argument arrays."""
        F = len(ctxs)
        L = ctxs[0].L
        hs = (C.c_void_p * F)(*[c.h for c in ctxs])
        cp = (C.c_void_p * F)(*[int(p) for p in cones_dev_ptrs])
        ip = (C.c_void_p * F)(*[int(p) for p in idx_dev_ptrs])
        na = (C.c_int * F)(*[int(v) for v in ns])
        pa = _f64(np.asarray(poses, dtype=np.float64).reshape(F, 3))
        thr_c = C.c_double(thr)

        def launch():
            rc = L.slam_b200_assoc_bulk_frames_dev(F, hs, cp, na, _dp(pa), thr_c, gate, algo, ip)
            if rc < 0:
                raise SlamB200Error(f"assoc_bulk_frames_dev failed ({rc}): {L.slam_b200_last_error(ctxs[0].h)}")
            return rc
        launch.keep = (hs, cp, ip, na, pa)
        return launch

    # ---- graph -----------------------------------------------------------------------------------
    def graph_clear(self):
        self._ck(self.L.slam_b200_graph_clear(self.h))

    def graph_add_pose(self, vid, x, y, th):
        return self._ck(self.L.slam_b200_graph_add_pose(self.h, int(vid), float(x), float(y), float(th)), "add_pose")

    def graph_add_landmark(self, vid, x, y):
        return self._ck(self.L.slam_b200_graph_add_landmark(self.h, int(vid), float(x), float(y)), "add_landmark")

    def graph_add_edge_se2(self, a, b, z, info):
        z = _f64(z); info = _f64(info)
        return self._ck(self.L.slam_b200_graph_add_edge_se2(self.h, int(a), int(b), _dp(z), _dp(info)), "add_edge_se2")

    def graph_add_odometry(self, a, b, pose, info):
        pose = _f64(pose); info = _f64(info)
        return self._ck(self.L.slam_b200_graph_add_odometry(self.h, int(a), int(b), _dp(pose), _dp(info)), "add_odometry")

    def graph_add_edge_se2_xy(self, p, l, z, info):
        z = _f64(z); info = _f64(info)
        return self._ck(self.L.slam_b200_graph_add_edge_se2_xy(self.h, int(p), int(l), _dp(z), _dp(info)), "add_edge_se2_xy")

    def graph_set_fixed(self, vid, flag=True):
        return self._ck(self.L.slam_b200_graph_set_fixed(self.h, int(vid), int(bool(flag))), "set_fixed")

    def graph_load(self, g):
        a = dict(pose_ids=_i32(g.pose_ids), pose_est=_f64(g.pose_est), lm_ids=_i32(g.lm_ids), lm_est=_f64(g.lm_est),
                 eo_from=_i32(g.eo_from), eo_to=_i32(g.eo_to), eo_z=_f64(g.eo_z), eo_info=_f64(g.eo_info),
                 el_pose=_i32(g.el_pose), el_lm=_i32(g.el_lm), el_z=_f64(g.el_z), el_info=_f64(g.el_info),
                 fixed=_i32(g.fixed_ids))
        self._ck(self.L.slam_b200_graph_load(
            self.h, len(a["pose_ids"]), _ip(a["pose_ids"]), _dp(a["pose_est"]), len(a["lm_ids"]), _ip(a["lm_ids"]),
            _dp(a["lm_est"]), len(a["eo_from"]), _ip(a["eo_from"]), _ip(a["eo_to"]), _dp(a["eo_z"]), _dp(a["eo_info"]),
            len(a["el_pose"]), _ip(a["el_pose"]), _ip(a["el_lm"]), _dp(a["el_z"]), _dp(a["el_info"]),
            len(a["fixed"]), _ip(a["fixed"])), "graph_load")

    def graph_set_values(self, pose_est=None, lm_est=None, eo_z=None, el_z=None):
        pe = None if pose_est is None else _f64(pose_est); le = None if lm_est is None else _f64(lm_est)
        oz = None if eo_z is None else _f64(eo_z); lz = None if el_z is None else _f64(el_z)
        self._ck(self.L.slam_b200_graph_set_values(self.h, _dp(pe), _dp(le), _dp(oz), _dp(lz)), "set_values")

    def graph_optimize(self, iters=10):
        chi2 = np.zeros(max(iters, 1))
        n = self._ck(self.L.slam_b200_graph_optimize(self.h, int(iters), _dp(chi2)), "graph_optimize") \
            if True else 0
        return n, chi2[:max(n, 0)]

    def graph_optimize_rc(self, iters=10):
        """Raw g2o-convention return code (-1 nothing to optimise, 0 failure, else iterations)."""
        chi2 = np.zeros(max(iters, 1))
        rc = self.L.slam_b200_graph_optimize(self.h, int(iters), _dp(chi2))
        if rc < -1:
            self._ck(rc, "graph_optimize")
        return rc, chi2[:max(rc, 0)]

    def graph_prepare(self):
        return self._ck(self.L.slam_b200_graph_prepare(self.h), "graph_prepare")

    def graph_prepare_assembly_only(self):
        return self._ck(self.L.slam_b200_graph_prepare_assembly_only(self.h), "graph_prepare_assembly_only")

    def graph_iterate_async(self, iters):
        self._ck(self.L.slam_b200_graph_iterate_async(self.h, int(iters)), "graph_iterate_async")

    def graph_finish(self, cap=64):
        chi2 = np.zeros(cap)
        rc = self.L.slam_b200_graph_finish(self.h, _dp(chi2), cap)
        if rc < -1:
            self._ck(rc, "graph_finish")
        return rc, chi2[:max(rc, 0)]

    def graph_reset_device(self):
        self._ck(self.L.slam_b200_graph_reset_device(self.h), "graph_reset_device")

    def graph_snapshot(self):
        self._ck(self.L.slam_b200_graph_snapshot(self.h), "graph_snapshot")

    def graph_restore_async(self):
        self._ck(self.L.slam_b200_graph_restore_async(self.h), "graph_restore_async")

    def profile_enable(self, on=True):
        self._ck(self.L.slam_b200_profile_enable(self.h, int(bool(on))), "profile_enable")

    def profile_read(self):
        out = np.zeros(8)
        self._ck(self.L.slam_b200_profile_read(self.h, _dp(out)), "profile_read")
        return dict(assemble_ms=out[0], factor_ms=out[1], forward_ms=out[2], backward_ms=out[3], update_ms=out[4],
                    iterations=int(out[5]))

    def fp64_peak_tflops(self):
        v = C.c_double(0)
        self._ck(self.L.slam_b200_fp64_peak(self.h, C.byref(v)), "fp64_peak")
        return v.value

    def graph_assemble_async(self, p0, p1):
        self._ck(self.L.slam_b200_graph_assemble_async(self.h, int(p0), int(p1)), "graph_assemble_async")

    def graph_shard_landmarks(self, p0, p1):
        l0 = C.c_int32(0); l1 = C.c_int32(0)
        self._ck(self.L.slam_b200_graph_shard_landmarks(self.h, int(p0), int(p1), C.byref(l0), C.byref(l1)), "shard_landmarks")
        return l0.value, l1.value

    def xchg_create(self, world, rank, cap):
        """Allocates this rank's peer-exchange region; returns its 64-byte CUDA IPC handle."""
        buf = (C.c_ubyte * 64)()
        self._ck(self.L.slam_b200_xchg_create(self.h, int(world), int(rank), int(cap), buf), "xchg_create")
        return bytes(buf)

    def xchg_connect(self, handles, ranges):
        """handles: world x 64 bytes (rank order); ranges: world x 2 int32 landmark ranges."""
        hb = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        rg = _i32(np.asarray(ranges).reshape(-1))
        self._ck(self.L.slam_b200_xchg_connect(self.h, hb, _ip(rg)), "xchg_connect")

    def graph_assemble_exchange_async(self, p0, p1):
        self._ck(self.L.slam_b200_graph_assemble_exchange_async(self.h, int(p0), int(p1)), "graph_assemble_exchange_async")

    def debug_guard_check(self):
        """(band bytes changed, arrays checked); (-1, 0) when SLAM_B200_GUARD is off."""
        n = C.c_long(0)
        self.L.slam_b200_debug_guard_check.restype = C.c_long
        self.L.slam_b200_debug_guard_check.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        bad = self.L.slam_b200_debug_guard_check(self.h, C.byref(n))
        return int(bad), int(n.value)

    def xchg_set_timeout_ms(self, ms):
        self._ck(self.L.slam_b200_xchg_set_timeout_ms(self.h, float(ms)), "xchg_set_timeout_ms")

    def xchg_error(self):
        return self._ck(self.L.slam_b200_xchg_error(self.h), "xchg_error")

    def graph_solve_async(self):
        self._ck(self.L.slam_b200_graph_solve_async(self.h), "graph_solve_async")

    def graph_system_dev(self, which):
        p = c_dp()
        n = self.L.slam_b200_graph_system_dev(self.h, int(which), C.byref(p))
        self._ck(int(n), "graph_system_dev")
        return C.cast(p, C.c_void_p).value, int(n)

    def graph_chi2(self):
        v = C.c_double(0)
        self._ck(self.L.slam_b200_graph_chi2(self.h, C.byref(v)), "graph_chi2")
        return v.value

    def graph_get_vertex(self, vid):
        out = np.zeros(3)
        d = self._ck(self.L.slam_b200_graph_get_vertex(self.h, int(vid), _dp(out)), "get_vertex")
        return out[:d]

    def graph_get_estimates(self):
        P = self.L.slam_b200_graph_num_poses(self.h); Ln = self.L.slam_b200_graph_num_landmarks(self.h)
        pe = np.zeros((max(P, 1), 3)); le = np.zeros((max(Ln, 1), 2))
        self._ck(self.L.slam_b200_graph_get_estimates(self.h, _dp(pe), _dp(le)), "get_estimates")
        return pe[:P], le[:Ln]

    def graph_export_system(self):
        n = C.c_int(0)
        nnz = self._ck(int(self.L.slam_b200_graph_export_system(self.h, C.byref(n), None, None, None, None)), "export")
        Ap = np.zeros(n.value + 1, dtype=np.int32); Ai = np.zeros(max(nnz, 1), dtype=np.int32)
        Ax = np.zeros(max(nnz, 1)); b = np.zeros(max(n.value, 1))
        self._ck(int(self.L.slam_b200_graph_export_system(self.h, C.byref(n), _ip(Ap), _ip(Ai), _dp(Ax), _dp(b))), "export")
        return dict(n=n.value, Ap=Ap, Ai=Ai[:nnz], Ax=Ax[:nnz], b=b[:n.value])

    def graph_stats(self):
        out = np.zeros(16)
        self._ck(self.L.slam_b200_graph_stats(self.h, _dp(out)), "graph_stats")
        keys = ["n", "n_blocks", "n_fronts", "n_levels", "nnz_H_upper", "nnz_L", "factor_flops", "max_front",
                "front_storage", "symbolic_seconds", "upload_seconds", "nV", "nFbig", "n_offdiag_blocks",
                "structure_seconds", "launch_lists_seconds"]
        return dict(zip(keys, out.tolist()))

    # ---- batched replicas ------------------------------------------------------------------------
    def batch_upload(self, pose_est, lm_est, eo_z, el_z):
        pe = _f64(pose_est); le = _f64(lm_est); oz = _f64(eo_z); lz = _f64(el_z)
        self._ck(self.L.slam_b200_batch_upload(self.h, pe.shape[0], _dp(pe), _dp(le), _dp(oz), _dp(lz)), "batch_upload")

    def batch_iterate_async(self, iters):
        self._ck(self.L.slam_b200_batch_iterate_async(self.h, int(iters)), "batch_iterate_async")

    def batch_download(self, R, P, Ln, iters):
        pe = np.zeros((R, P, 3)); le = np.zeros((R, Ln, 2)); chi2 = np.zeros((R, max(iters, 1)))
        done = np.zeros(R, dtype=np.int32)
        self._ck(self.L.slam_b200_batch_download(self.h, _dp(pe), _dp(le), _dp(chi2), max(iters, 1), _ip(done)),
                 "batch_download")
        return pe, le, chi2, done

    def graph_optimize_batch(self, pose_est, lm_est, eo_z, el_z, iters=10):
        pe = _f64(pose_est).copy(); le = _f64(lm_est).copy(); oz = _f64(eo_z); lz = _f64(el_z)
        R = pe.shape[0]
        chi2 = np.zeros((R, max(iters, 1))); done = np.zeros(R, dtype=np.int32)
        self._ck(self.L.slam_b200_graph_optimize_batch(self.h, R, _dp(pe), _dp(le), _dp(oz), _dp(lz), int(iters),
                                                       _dp(chi2), _ip(done)), "graph_optimize_batch")
        return pe, le, chi2, done


class SymbolicAnalysis:
    """Host-only symbolic analysis of a block pattern (no GPU needed)."""
    KEYS = {1: "pos", 2: "boff", 3: "level_ptr", 4: "piv0", 5: "npiv", 6: "nupd", 7: "parent", 8: "rows_ptr",
            9: "upd_rows", 10: "rel", 11: "asm_ptr", 12: "asm", 14: "dim", 15: "child_ptr", 16: "children"}

    def __init__(self, dims, pair_a, pair_b, leaf_size=8):
        L = lib()
        dims = _i32(dims); a = _i32(pair_a); b = _i32(pair_b)
        h = L.slam_b200_symbolic_create(len(dims), _ip(dims), len(a), _ip(a), _ip(b), int(leaf_size))
        if not h:
            raise SlamB200Error("symbolic_create failed")
        h = C.c_void_p(h)
        try:
            def get(what):
                n = L.slam_b200_symbolic_export(h, what, None, 0)
                out = np.zeros(max(n, 1), dtype=np.int32)
                L.slam_b200_symbolic_export(h, what, _ip(out), n)
                return out[:n]
            sz = get(0)
            self.nb, self.n, self.nf, self.nlevels = (int(v) for v in sz[:4])
            self.max_front = int(sz[6])
            for w, k in self.KEYS.items():
                setattr(self, k, get(w))
            self.asm = self.asm.reshape(-1, 4)
            self.nnzL = L.slam_b200_symbolic_stat(h, 0)
            self.flops = L.slam_b200_symbolic_stat(h, 1)
            self.seconds = L.slam_b200_symbolic_stat(h, 3)
            self.front_storage = L.slam_b200_symbolic_stat(h, 4)
            # H layout the analysis assumed
            hd = np.concatenate([[0], np.cumsum(dims.astype(np.int64) ** 2)])
            self.hoff_diag = hd[:-1]
            ho = hd[-1] + np.concatenate([[0], np.cumsum(dims[a].astype(np.int64) * dims[b])])
            self.hoff_off = ho[:-1]
            self.nH = int(ho[-1])
            # host plan of the tiled batched factorisation (csrc/tileplan.h)

            def tget(what):
                n = L.slam_b200_symbolic_tileplan(h, what, None, 0)
                out = np.zeros(max(n, 1), dtype=np.int32)
                L.slam_b200_symbolic_tileplan(h, what, _ip(out), n)
                return out[:n]
            info = tget(0)
            self.tile_ok = bool(info[0])
            if self.tile_ok:
                self.tile_max_T, self.tile_nF, self.tile_rhs_base = int(info[2]), int(info[4]), int(info[5])
                self.tile_T, self.tile_KT, self.tile_fptr = tget(1), tget(2), tget(3)
                self.tile_item_ptr, self.tile_item_nv = tget(4), tget(5)
                self.tile_items = tget(6).reshape(-1, 2)
        finally:
            L.slam_b200_symbolic_destroy(h)

"""Body of test_slam_host_gpu.py::test_localizer_repair_equals_the_oracle, run as its own process:
the drop-in Slam with the opt-in localiser repair against the oracle, frame by frame.  Exit code 0 = equal."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from conftest import load_pkg  # noqa: E402

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


def main(window):
    pkg = load_pkg()
    pkg.build()
    from importlib import import_module
    from oracle import oracle
    orc = oracle.load("best")
    synth = pkg.synth
    L = C.CDLL(import_module(pkg.__name__ + "._build").HOSTLIB)
    L.slamhost_create.restype = C.c_void_p
    L.slamhost_create.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
    L.slamhost_destroy.argtypes = [C.c_void_p]
    L.slamhost_set_localizer_repair.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.slamhost_perform.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, C.c_float, C.c_double, c_ip, c_ip]
    L.slamhost_last_error.restype = C.c_char_p
    L.slamhost_state.argtypes = [C.c_void_p, c_ip]
    L.slamhost_chi2_log.argtypes = [C.c_void_p, c_dp]
    L.slamhost_draw_current_pose.argtypes = [C.c_void_p, c_dp]
    L.slamhost_draw_cones.argtypes = [C.c_void_p, c_dp, c_dp, c_ip, c_ip]

    d = synth.trackdrive(2, poses_per_lap=400, seed=21)
    h = C.c_void_p(L.slamhost_create(synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD, 20, 0))
    assert h, L.slamhost_last_error()
    L.slamhost_set_localizer_repair(h, 1, window)
    o = orc.slam(synth.SAME_CONE_THRESHOLD, synth.CONE_MAPPING_THRESHOLD)
    o.set_localizer_repair(True, window)
    n_loc = 0
    worst = 0.0
    for k, (fr, p) in enumerate(zip(d.frames, d.poses_noisy)):
        fr = np.asfortranarray(fr, dtype=np.float64); p = np.ascontiguousarray(p, dtype=np.float64)
        n = fr.shape[1]
        idx = np.zeros(max(n, 1), dtype=np.int32); st = np.zeros(max(n, 1), dtype=np.int32)
        rc = L.slamhost_perform(h, fr.ctypes.data_as(c_dp), n, p.ctypes.data_as(c_dp), 0.0, 0.0,
                                idx.ctypes.data_as(c_ip), st.ctypes.data_as(c_ip))
        assert rc != -100, L.slamhost_last_error()
        rco, idxo, sto = o.perform(fr, p)
        assert rc == rco, (k, rc, rco)
        assert np.array_equal(idx[:n], idxo) and np.array_equal(st[:n], sto), k
        if rc == 2:
            n_loc += 1
            e = np.zeros(3)
            L.slamhost_draw_current_pose(h, e.ctypes.data_as(c_dp))
            eo = o.send_pose()
            err = max(np.max(np.abs(e[:2] - eo[:2])) / 100.0, abs(e[2] - eo[2]))
            worst = max(worst, err)
            assert err <= 1e-6, (k, e, eo)      # 1e-6 relative to the 100 m scale of the track
    assert n_loc > 300
    sd = np.zeros(8, dtype=np.int32)
    L.slamhost_state(h, sd.ctypes.data_as(c_ip))
    so = o.state()
    assert sd[4] == so["optimize_calls"] > 250 and sd[5] == so["last_iterations"] == 10, (sd, so)
    chi2 = np.zeros(max(int(sd[6]), 1))
    L.slamhost_chi2_log(h, chi2.ctypes.data_as(c_dp))
    assert np.allclose(chi2[:sd[6]], o.chi2_log(), rtol=1e-6, atol=1e-12)
    M = int(sd[7])
    x = np.zeros(M); y = np.zeros(M); t = np.zeros(M, dtype=np.int32); ids = np.zeros(M, dtype=np.int32)
    L.slamhost_draw_cones(h, x.ctypes.data_as(c_dp), y.ctypes.data_as(c_dp), t.ctypes.data_as(c_ip), ids.ctypes.data_as(c_ip))
    ox, oy, ot = o.map()
    assert np.array_equal(t, ot) and np.max(np.abs(x - ox)) <= 1e-4 and np.max(np.abs(y - oy)) <= 1e-4
    L.slamhost_destroy(h)
    print(f"localiser repair, window {window}: {n_loc} localiser frames, {sd[4]} optimise calls, "
          f"worst sent-pose deviation {worst:.3g} (relative)")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10)

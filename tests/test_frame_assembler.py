"""the deterministic, clock-injectable frame assembler (csrc/host/frame_assembler.cpp)
against the semantics of Slam::nextCone / initializeCollection / isKeyframe (src/slam.cpp:67-152, 221-257,
286-295).  Pure host logic: runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import skip_if_sanitizer_runtime_unusable

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def host(pkg):
    from importlib import import_module
    b = import_module(pkg.__name__ + "._build")
    pkg.build()
    L = C.CDLL(b.HOSTLIB)
    L.frameasm_create.restype = C.c_void_p
    L.frameasm_create.argtypes = [C.c_int, C.c_double]
    L.frameasm_destroy.argtypes = [C.c_void_p]
    L.frameasm_add_direction.argtypes = [C.c_void_p, C.c_uint, C.c_float, C.c_float, C.c_longlong]
    L.frameasm_add_distance.argtypes = [C.c_void_p, C.c_uint, C.c_float, C.c_longlong]
    L.frameasm_add_type.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_longlong]
    L.frameasm_poll.argtypes = [C.c_void_p, C.c_longlong, c_dp, C.c_int]
    L.frameasm_state.argtypes = [C.c_void_p, c_ip]
    return L


class Asm:
    def __init__(self, L, gather_ms=110, between=0.5):
        self.L = L
        self.h = C.c_void_p(L.frameasm_create(gather_ms, between))

    def send_frame(self, cols, t0, dt=100):
        """cols: (4,N) az, zen, range, type; messages arrive dt microseconds apart from t0 on."""
        t = t0
        for i in range(cols.shape[1]):
            self.L.frameasm_add_direction(self.h, i, cols[0, i], cols[1, i], t); t += dt
            self.L.frameasm_add_distance(self.h, i, cols[2, i], t); t += dt
            self.L.frameasm_add_type(self.h, i, int(cols[3, i]), t); t += dt
        return t

    def poll(self, now):
        out = np.zeros((1000, 4))
        n = self.L.frameasm_poll(self.h, now, out.ctypes.data_as(c_dp), 1000)
        return None if n < 0 else out[:n].T.copy()

    def state(self):
        s = np.zeros(5, dtype=np.int32)
        self.L.frameasm_state(self.h, s.ctypes.data_as(c_ip))
        return dict(zip(["open", "gathered", "dropped", "out_of_range", "capacity"], s.tolist()))


def test_frame_is_gathered_after_the_window_and_matches_the_messages(host, c1_drive):
    a = Asm(host)
    fr = c1_drive.frames[5]
    t_end = a.send_frame(fr, 1_000_000)
    assert a.state()["open"] == 1
    assert a.poll(1_000_000 + 110_000) is None            # elapsed must EXCEED the window (slam.cpp:231)
    got = a.poll(1_000_000 + 110_001)
    assert got is not None and got.shape == fr.shape
    assert np.array_equal(got, fr)                         # float32 wire fields widened: exact
    st = a.state()
    assert st["open"] == 0 and st["gathered"] == 1 and st["capacity"] == 1000   # 4x100 -> 4x1000 (slam.cpp:46,244)
    assert a.poll(2_000_000) is None                       # nothing open


def test_missing_ids_leave_zero_columns_and_window_starts_at_first_message(host):
    a = Asm(host)
    host.frameasm_add_distance(a.h, 3, 7.5, 500)           # first message opens the frame at t=500
    host.frameasm_add_direction(a.h, 1, 10.0, 0.0, 90_000)
    host.frameasm_add_type(a.h, 1, 2, 100_000)
    assert a.poll(110_400) is None
    got = a.poll(110_501)
    assert got.shape == (4, 4)                             # leftCols(lastObjectId + 1)
    assert np.array_equal(got[:, 0], [0, 0, 0, 0]) and np.array_equal(got[:, 2], [0, 0, 0, 0])
    assert np.array_equal(got[:, 1], [10.0, 0.0, 0.0, 2.0]) and np.array_equal(got[:, 3], [0, 0, 7.5, 0])


def test_keyframe_gate_uses_milliseconds_against_the_raw_setting(host):
    """isKeyframe compares |delta| in ms with timeBetweenKeyframes as is (slam.cpp:288-290)."""
    a = Asm(host, gather_ms=1, between=500.0)              # 500 'units' == 500 ms in the reference's arithmetic
    cols = np.array([[5.0], [0.0], [3.0], [1.0]])
    t = 10_000_000
    a.send_frame(cols, t)
    assert a.poll(t + 2_000) is not None                   # first frame: keyframe stamp is zero -> always passes
    a.send_frame(cols, t + 100_000)
    assert a.poll(t + 102_000) is None                     # 100 ms later: gathered but dropped by the gate
    assert a.state()["dropped"] == 1
    a.send_frame(cols, t + 600_000)
    assert a.poll(t + 602_000 + 1_000) is not None         # > 500 ms after the last keyframe


def test_object_ids_beyond_the_collector_are_dropped_not_written(host):
    a = Asm(host)
    host.frameasm_add_direction(a.h, 100, 1.0, 0.0, 0)     # capacity is 100 before the first frame
    assert a.state()["out_of_range"] == 1 and a.state()["open"] == 0
    host.frameasm_add_direction(a.h, 99, 1.0, 0.0, 10)
    got = a.poll(10 + 110_001)
    assert got.shape == (4, 100)
    host.frameasm_add_direction(a.h, 999, 1.0, 0.0, 500_000)   # 4x1000 now
    assert a.state()["out_of_range"] == 1 and a.state()["open"] == 1


def test_replay_is_deterministic(host, c1_drive):
    outs = []
    for _ in range(2):
        a = Asm(host)
        t = 0
        frames = []
        for fr in c1_drive.frames[:50]:
            t_end = a.send_frame(fr, t)
            f = a.poll(t + 110_001)
            frames.append(f)
            t += 150_000
        outs.append(frames)
    for x, y, fr in zip(outs[0], outs[1], c1_drive.frames[:50]):
        assert x is not None and np.array_equal(x, y) and np.array_equal(x, fr)


def test_host_mirror_compiles_against_the_reference_eigen(pkg):
    """INTEGRATION.md claims the Slam/Cone bodies compile with the reference's own Eigen types."""
    import subprocess
    inc = "/root/reference/thirdparty"
    if not os.path.isdir(os.path.join(inc, "Eigen")):
        pytest.skip("reference tree absent")
    src = os.path.join(os.path.dirname(pkg.__file__), "csrc", "host")
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-DSLAM_B200_WITH_EIGEN", "-I" + inc,
                        os.path.join(src, "slam.cpp"), os.path.join(src, "cone.cpp"),
                        os.path.join(src, "frame_assembler.cpp"), os.path.join(src, "slam_c.cpp")],
                       capture_output=True, text=True)
    skip_if_sanitizer_runtime_unusable(r.stdout + r.stderr)
    assert r.returncode == 0, r.stderr


def test_slam_lock_discipline_under_tsan(pkg, tmp_path):
    """The drop-in Slam class (csrc/host/slam.cpp) under ThreadSanitizer: a frame thread running performSLAM through
    mapping, loop closure and localiser frames (reference mode and opt-in repair) while a viewer thread calls draw* /
    buildConePacket, over a stub of the sixteen C-ABI calls it makes (tests/tsan_slam_stub_backend.cpp -- canned
    records, test infrastructure only).  The reference's own discipline fails this: addPoseToGraph grows
    m_connectivityGraph under the optimizer mutex while drawGraph copies it under map + sensor (slam.cpp:440, 780-784),
    and localizer -> sendCones takes optimizer -> map against addConesToMap's map -> optimizer."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    host = os.path.join(os.path.dirname(here), pkg.__name__, "csrc", "host")
    exe = str(tmp_path / "tsan_slam")
    cmd = ["g++", "-std=c++14", "-O1", "-g", "-fsanitize=thread", "-I" + host, os.path.join(here, "tsan_slam_driver.cpp"),
           os.path.join(here, "tsan_slam_stub_backend.cpp"), os.path.join(host, "slam.cpp"), os.path.join(host, "cone.cpp"),
           "-o", exe, "-lpthread"]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0:
        pytest.skip("sanitizer runtime not available: " + b.stderr[-300:])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=240, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
    skip_if_sanitizer_runtime_unusable(r.stdout + r.stderr)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout[-500:] + r.stderr[-3000:]

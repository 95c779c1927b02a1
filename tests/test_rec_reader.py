"""Recording reader + deterministic replay of the front half (csrc/host/rec_reader.cpp).

tests/golden/c1_head.rec was written by THE REFERENCE'S OWN cluon (serializeEnvelope + the message set generated
from the reference's .odvd; tests/golden/make_rec_golden.py), so the wire format is pinned by reference code:
the reader has to return every field the generator put in, bit for bit (floats are float32 on the wire)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import skip_if_sanitizer_runtime_unusable

HERE = os.path.dirname(os.path.abspath(__file__))
REC = os.path.join(HERE, "golden", "c1_head.rec")
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_rec_golden as gen  # noqa: E402

KIND_ID = {"W": 19, "Y": 1031, "H": 1051, "G": 1116, "T": 1131, "D": 1133, "R": 1134, "X": 1046}
c_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hostlib(pkg):
    from importlib import import_module
    b = import_module(pkg.__name__ + "._build")
    pkg.build()
    L = C.CDLL(b.HOSTLIB)
    i32p, u32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)
    L.slamrec_read.argtypes = [C.c_char_p, C.c_int, i32p, u32p, i64p, i64p, u32p, c_dp, i64p]
    L.slamrec_replay.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                 i32p, c_dp, c_dp, C.POINTER(C.c_float), c_dp, i64p, i64p]
    L.slamhost_wgs84_to_cartesian.argtypes = [c_dp, c_dp, c_dp]
    L.slamhost_heading_from_north.restype = C.c_double
    L.slamhost_heading_from_north.argtypes = [C.c_float]
    return L


def _read(L, path, cap=4096):
    dt = np.zeros(cap, dtype=np.int32); sd = np.zeros(cap, dtype=np.uint32); ts = np.zeros(cap, dtype=np.int64)
    sent = np.zeros(cap, dtype=np.int64); oid = np.zeros(cap, dtype=np.uint32); f = np.zeros((cap, 4)); st = np.zeros(3, dtype=np.int64)
    n = L.slamrec_read(path.encode(), cap, dt.ctypes.data_as(C.POINTER(C.c_int32)), sd.ctypes.data_as(C.POINTER(C.c_uint32)),
                       ts.ctypes.data_as(C.POINTER(C.c_int64)), sent.ctypes.data_as(C.POINTER(C.c_int64)),
                       oid.ctypes.data_as(C.POINTER(C.c_uint32)), f.ctypes.data_as(c_dp), st.ctypes.data_as(C.POINTER(C.c_int64)))
    return n, dt[:n], sd[:n], ts[:n], sent[:n], oid[:n], f[:n], st


def _replay(L, path, gather_ms=10, between=0.5, cap_frames=256, cap_cols=8192):
    nc = np.zeros(cap_frames, dtype=np.int32); cones = np.zeros(4 * cap_cols); odo = np.zeros((cap_frames, 3))
    yaw = np.zeros(cap_frames, dtype=np.float32); el = np.zeros(cap_frames); tu = np.zeros(cap_frames, dtype=np.int64)
    st = np.zeros(9, dtype=np.int64)
    n = L.slamrec_replay(path.encode(), gen.DETECT_CONE_ID, gen.ESTIMATION_ID, gather_ms, between, gen.REF_LAT, gen.REF_LON,
                         cap_frames, cap_cols, nc.ctypes.data_as(C.POINTER(C.c_int32)), cones.ctypes.data_as(c_dp), odo.ctypes.data_as(c_dp),
                         yaw.ctypes.data_as(C.POINTER(C.c_float)), el.ctypes.data_as(c_dp), tu.ctypes.data_as(C.POINTER(C.c_int64)),
                         st.ctypes.data_as(C.POINTER(C.c_int64)))
    frames, off = [], 0
    for k in range(n):
        frames.append(cones[4 * off:4 * (off + nc[k])].reshape(nc[k], 4).T.copy())
        off += nc[k]
    return frames, odo[:n], yaw[:n], el[:n], tu[:n], st


def test_every_envelope_cluon_wrote_is_read_back_bit_for_bit(hostlib, synth):
    msgs, _, _ = gen.messages(synth)
    n, dt, sd, ts, sent, oid, f, st = _read(hostlib, REC)
    assert n == len(msgs) == 1216 and st[0] == 0 and st[1] == 0 and st[2] == n
    f32 = lambda v: float(np.float32(v))
    for k, m in enumerate(msgs):
        kind, t, sender, a, b, c, d = m
        assert dt[k] == KIND_ID[kind] and sd[k] == sender and ts[k] == t and sent[k] == t, (k, m)
        if kind == "D":
            assert oid[k] == int(a) and f[k, 0] == f32(b) and f[k, 1] == f32(c)
        elif kind == "R":
            assert oid[k] == int(a) and f[k, 0] == f32(b)
        elif kind == "T":
            assert oid[k] == int(a) and f[k, 0] == float(int(b))
        elif kind == "G":
            assert f[k, 0] == a and f[k, 1] == b and f[k, 2] == f32(c) and f[k, 3] == f32(d)    # doubles exact
        elif kind == "W":
            assert f[k, 0] == a and f[k, 1] == b
        elif kind == "H":
            assert f[k, 0] == f32(a)
        elif kind == "Y":
            assert f[k, 0] == f32(a) and f[k, 1] == f32(b) and f[k, 2] == f32(c)


def test_replay_releases_the_frames_and_poses_that_were_recorded(hostlib, synth):
    """Front half on recorded time: cone messages of the cone sender are gathered per frame (interleaved
    direction / distance / type, scrambled object order), pose messages of the estimation sender set the
    odometry (Geolocation, or the split WGS84 + heading pair) and the yaw rate; other senders and other
    message types are ignored, like the data triggers in main()."""
    msgs, frames, geo = gen.messages(synth)
    got, odo, yaw, el, tu, st = _replay(hostlib, REC)
    assert st[0] == len(msgs) and st[5] == 0
    assert st[3] == 2 * len(frames) and st[4] == len(frames)           # wrong senders, ignored message type
    assert st[1] == 3 * sum(fr.shape[1] for fr in frames)
    assert len(got) == len(frames) == st[8] == st[6] and st[7] == 0
    ref = np.array([gen.REF_LAT, gen.REF_LON])
    for k, (fr, g) in enumerate(zip(frames, got)):
        want = fr.astype(np.float32).astype(np.float64)                # float32 on the wire (odvd:294-303)
        assert g.shape == want.shape and np.array_equal(g, want), k
        lat, lon, heading, split = geo[k]
        xy = np.zeros(2); pos = np.array([lat, lon])
        hostlib.slamhost_wgs84_to_cartesian(ref.ctypes.data_as(c_dp), pos.ctypes.data_as(c_dp), xy.ctypes.data_as(c_dp))
        assert odo[k, 0] == xy[0] and odo[k, 1] == xy[1]
        if split:
            assert odo[k, 2] == hostlib.slamhost_heading_from_north(float(np.float32(heading + 3.14159265)))
        else:
            assert odo[k, 2] == float(np.float32(heading))
        assert yaw[k] == np.float32(np.float32(0.2 * np.sin(k)) / np.float32(4))
    assert np.all(np.diff(tu) > 0)


def test_keyframe_gate_and_gathering_window_on_recorded_time(hostlib, synth):
    """A window longer than the frame spacing merges consecutive frames; a keyframe interval longer than the
    spacing drops frames (in the reference's milliseconds-vs-raw-setting units, slam.cpp:286-295)."""
    _, frames, _ = gen.messages(synth)
    got, *_, st = _replay(hostlib, REC, gather_ms=10, between=700.0)   # 700 "ms" > 600 ms spacing: every other frame
    # frames 0, 2, 4, ... pass the gate; the frame still open at the end of the file is released 'long after'
    assert st[6] == len(frames) and len(got) == st[8] == len(frames) // 2 + 1 and st[7] == len(frames) - len(got)
    got2, *_, st2 = _replay(hostlib, REC, gather_ms=900)                # 0.9 s window swallows the next frame's messages
    assert st2[6] < len(frames) and len(got2) == st2[8]


def test_corrupt_and_truncated_streams(hostlib, tmp_path):
    data = open(REC, "rb").read()
    n0 = _read(hostlib, REC)[0]
    p = tmp_path / "garbage_prefix.rec"
    p.write_bytes(b"\x00\x01\x0d\x02garbage" + data)                     # resync on the 0x0D 0xA4 header
    n, *_, st = _read(hostlib, str(p))
    assert n == n0 and st[0] == 11
    p = tmp_path / "truncated.rec"
    p.write_bytes(data[:-7])                                              # last envelope incomplete
    n, *_, st = _read(hostlib, str(p))
    assert n == n0 - 1 and st[1] > 0
    p = tmp_path / "empty.rec"
    p.write_bytes(b"")
    assert _read(hostlib, str(p))[0] == 0
    assert hostlib.slamrec_read(b"/nonexistent/file.rec", 0, None, None, None, None, None, None, None) == -1


def test_mutated_streams_under_sanitizers(pkg, tmp_path):
    """Reader + replay on 3,000 corrupted variants of the golden recording under AddressSanitizer and UBSan:
    no out-of-bounds read, no overflow, no hang, whatever the 24-bit length, varints and nested lengths say."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    host = os.path.join(os.path.dirname(here), pkg.__name__, "csrc", "host")
    exe = str(tmp_path / "fuzz_rec_driver")
    cmd = ["g++", "-std=c++14", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I" + host,
           os.path.join(here, "fuzz_rec_driver.cpp")] + [os.path.join(host, f) for f in ("rec_reader.cpp", "frame_assembler.cpp", "wgs84.cpp")] + ["-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0:
        pytest.skip("sanitizer runtime not available: " + b.stderr[-300:])
    r = subprocess.run([exe, os.path.join(here, "golden", "c1_head.rec"), "7", "3000"], capture_output=True, text=True, timeout=600)
    skip_if_sanitizer_runtime_unusable(r.stdout + r.stderr)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout[-500:] + r.stderr[-3000:]


@pytest.mark.skipif(not os.path.exists("/root/reference/src/cluon-complete-build.hpp"), reason="reference tree not on this machine")
def test_rec_fixture_is_what_the_reference_cluon_writes_today():
    committed = open(REC, "rb").read()
    try:
        subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_rec_golden.py")], check=True, capture_output=True, timeout=900)
        assert open(REC, "rb").read() == committed
    finally:
        open(REC, "wb").write(committed)

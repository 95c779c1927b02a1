"""The drop-in boundary inside the reference tree -- CPU part.

integration/slam_b200.patch swaps the back half of the reference's class Slam for calls into the C ABI;
integration/build.sh applies it to a scratch copy of /root/reference/src, compiles the patched tree against
libslam_b200.so and, with the same harness, the unmodified reference over the g2o facade.  The harness
(integration/public_api_harness.cpp) drives the class through its PUBLIC interface only: OD4 session constructor,
nextPose / nextYawRate / nextCone with cluon Envelopes, draw* to read the result.  tests/golden/public_api_reference.npz
is what the UNMODIFIED reference answers (tests/golden/make_public_api_golden.py).

Here (no GPU): the committed patch is current and applies; the reference still answers the golden file; and the golden
file pins the deterministic frame assembler (csrc/host/frame_assembler.cpp) + the oracle's back half against the
reference's own nextCone / initializeCollection / isKeyframe / performSLAM: the same messages, assembled on injected
time stamps instead of the wall clock, give the same association rows and map.  The GPU part is tests/test_integration_gpu.py."""
import ctypes as C
import os
import shutil
import struct
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
REF = "/root/reference"
sys.path.insert(0, GOLD)

have_ref = os.path.exists(os.path.join(REF, "src", "slam.cpp"))


def read_drive(path):
    b = open(path, "rb").read()
    n, thr, map_thr = struct.unpack_from("<idd", b, 0)
    assert n < 0
    off, frames = 20, []
    for _ in range(-n):
        yaw, el = struct.unpack_from("<fq", b, off); off += 12
        pose = np.frombuffer(b, dtype=np.float64, count=3, offset=off).copy(); off += 24
        (N,) = struct.unpack_from("<i", b, off); off += 4
        cols = np.frombuffer(b, dtype=np.float64, count=4 * N, offset=off).reshape(4, N, order="F").copy(); off += 32 * N
        frames.append((yaw, el, pose, cols))
    return thr, map_thr, frames


@pytest.mark.skipif(not have_ref, reason="needs the reference tree")
def test_patch_is_current_and_applies(tmp_path):
    out = tmp_path / "tree"
    patch = tmp_path / "regenerated.patch"
    subprocess.run([sys.executable, os.path.join(ROOT, "integration", "make_patched_tree.py"), REF, str(out),
                    "--write-patch", str(patch)], check=True, stdout=subprocess.DEVNULL)
    committed = open(os.path.join(ROOT, "integration", "slam_b200.patch")).read()
    assert committed == open(patch).read(), "integration/slam_b200.patch is stale: rerun integration/make_patched_tree.py"
    scratch = tmp_path / "scratch" / "src"
    os.makedirs(scratch)
    for f in ("slam.hpp", "slam.cpp"):
        shutil.copy(os.path.join(REF, "src", f), scratch / f)
    subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", os.path.join(ROOT, "integration", "slam_b200.patch")],
                   cwd=tmp_path / "scratch", check=True, stdout=subprocess.DEVNULL)
    for f in ("slam.hpp", "slam.cpp"):
        assert open(scratch / f).read() == open(out / "src" / f).read()
    hdr_old = open(os.path.join(REF, "src", "slam.hpp")).read()
    hdr_new = open(scratch / "slam.hpp").read()
    pub = lambda h: h[h.index("public:"):h.index("private:", h.index("public:"))]
    assert pub(hdr_old) == pub(hdr_new)                       # slam.hpp:52-62 byte for byte
    src_new = open(scratch / "slam.cpp").read()
    assert "g2o" not in hdr_new.replace("replaces g2o::SparseOptimizer", "") and "g2o::" not in src_new
    # the front half and the senders are untouched: every line of them is still there
    old_lines = open(os.path.join(REF, "src", "slam.cpp")).read().split("\n")
    a = next(i for i, l in enumerate(old_lines) if l.startswith("void Slam::nextCone"))
    b = next(i for i, l in enumerate(old_lines) if l.startswith("void Slam::localizer"))
    assert "\n".join(old_lines[a:b]) in src_new


@pytest.mark.skipif(not have_ref, reason="needs the reference tree")
def test_reference_still_answers_the_golden_file():
    import make_public_api_golden as mk
    subprocess.run(["sh", os.path.join(ROOT, "integration", "build.sh")], check=True, stdout=subprocess.DEVNULL)
    g = np.load(os.path.join(GOLD, "public_api_reference.npz"))
    r = mk.run(os.path.join(ROOT, "integration", "_build", "ref_public_replay"), os.path.join(GOLD, "public_api_drive_odd.bin"))
    for k, v in r.items():
        assert np.array_equal(v, g["odd_" + k], equal_nan=True), k


def test_public_api_c1_equals_the_direct_replay():
    """Through Envelopes and the wall-clock front half the reference makes the same 8,381 association decisions as
    when performSLAM is called directly (c1_replay_reference.npz); only the float32 heading of the Geolocation
    message moves the optimised map, by about a micrometre."""
    g = np.load(os.path.join(GOLD, "public_api_reference.npz"))
    r = np.load(os.path.join(GOLD, "c1_replay_reference.npz"))
    assert np.array_equal(g["c1_graph_ptr"], r["row_ptr"]) and np.array_equal(g["c1_graph_ids"], r["row_ids"])
    assert len(g["c1_map"]) == len(r["map_x"]) == 300
    assert np.max(np.abs(g["c1_map"][:, 0] - r["map_x"])) < 1e-5 and np.max(np.abs(g["c1_map"][:, 1] - r["map_y"])) < 1e-5
    assert np.array_equal(g["c1_map"][:, 2], r["map_type"]) and np.array_equal(g["c1_map"][:, 3], r["map_id"])


@pytest.mark.parametrize("name", ["odd", "c1", "loop1", "loop2", "loop5", "nan0", "nan2", "yaw1"])
def test_frame_assembler_and_oracle_equal_the_reference_front_and_back_half(pkg, orc, name):
    """the messages the harness sent, assembled by FrameAssembler on injected time stamps and run
    through the oracle's performSLAM, against what the reference's own nextCone / initializeCollection / isKeyframe /
    performSLAM made of them on the wall clock."""
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
    L.frameasm_poll.argtypes = [C.c_void_p, C.c_longlong, C.POINTER(C.c_double), C.c_int]
    g = np.load(os.path.join(GOLD, "public_api_reference.npz"))
    thr, map_thr, frames = read_drive(os.path.join(GOLD, "public_api_drive_%s.bin" % name))
    gather_ms = 10
    h = C.c_void_p(L.frameasm_create(gather_ms, 0.5))
    s = orc.slam(thr, map_thr)
    gp, gi, gposes, gframes = g[name + "_graph_ptr"], g[name + "_graph_ids"], g[name + "_poses"], g[name + "_frames"]
    stored, t = 0, 1000 * 1000000
    try:
        for k, (yaw, el, pose, cols) in enumerate(frames):
            sent = 0
            for i in range(cols.shape[1]):
                c = cols[:, i]
                if not c.any():
                    continue                                  # absent objectId, like the harness
                L.frameasm_add_direction(h, i, c[0], c[1], t)
                L.frameasm_add_distance(h, i, c[2], t)
                L.frameasm_add_type(h, i, int(np.uint32(c[3])), t)
                sent += 1
            out = np.zeros((1000, 4))
            n = L.frameasm_poll(h, t + gather_ms * 1000 + 1, out.ctypes.data_as(C.POINTER(C.c_double)), 1000)
            t += 100000
            if sent == 0:
                assert n < 0
                continue
            assert n == cols.shape[1]                          # leftCols(lastObjectId + 1)
            frame = out[:n].T.copy()
            assert int(gframes[k, 0]) in (stored, stored + 1)
            if int(gframes[k, 0]) == stored:                   # the reference dropped it: behind the 200 m gate
                assert abs(pose[0]) > 200 or abs(pose[1]) > 200
                continue
            # x, y exactly as nextPose computed them (WGS84 round trip of the harness: nanometres off the input),
            # heading as the float32 of the Geolocation message, yaw rate / elapsed as nextYawRate + performSLAM see them
            xy = gposes[stored, :2]
            assert np.max(np.abs(xy - pose[:2])) < 1e-7
            p = np.array([xy[0], xy[1], float(np.float32(pose[2]))])
            s.perform(frame, p, yaw_rate=float(np.float32(np.float32(yaw) * np.float32(4)) / np.float32(4)),
                      time_elapsed=abs(float(el)) / 1000000)
            row = s.connectivity_row(stored)
            assert np.array_equal(row, gi[gp[stored]:gp[stored + 1]]), (name, k)
            stored += 1
    finally:
        L.frameasm_destroy(h)
    assert stored == len(gposes)
    assert np.max(np.abs(s.poses() - gposes)) == 0.0          # stored poses incl. the heading correction: bit for bit
    mx, my, mt = s.map()
    gm = g[name + "_map"]
    assert len(mx) == len(gm) and np.array_equal(mt, gm[:, 2])
    assert np.array_equal(np.isnan(mx), np.isnan(gm[:, 0]))
    ok = ~np.isnan(mx)
    assert np.max(np.abs(mx[ok] - gm[ok, 0])) < 1e-9 and np.max(np.abs(my[ok] - gm[ok, 1])) < 1e-9

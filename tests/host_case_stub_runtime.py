"""Body of test_abi.py::test_host_code_over_a_stub_runtime (own process, LD_PRELOAD=<stub_cudart.so>): drives the
host code behind the C ABI on a box without a GPU -- see tests/stub_cudart.cpp.  Prints one JSON line.
SLAM_LIB overrides the library (profiles/tools/host_sanitize.sh points it at an ASan build)."""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from conftest import load_pkg, small_graph  # noqa: E402

pkg = load_pkg()
if os.environ.get("SLAM_LIB"):
    importlib.import_module(pkg.__name__ + "._build").LIB = os.environ["SLAM_LIB"]
stub = C.CDLL(os.environ["SLAM_STUB_CUDART"])
stub.stub_h2d_hash.restype = C.c_ulong
stub.stub_h2d_bytes.restype = C.c_size_t
synth = pkg.synth
ctx = pkg.Context(0)
L = ctx.L
c_dp = C.POINTER(C.c_double)
out = {"uploads": {}, "rc": {}}


def prepare(g, label):
    stub.stub_reset()
    ctx.graph_load(g)
    n = ctx.graph_prepare()
    st = ctx.graph_stats()
    out["uploads"][label] = {"n": int(n), "bytes": int(stub.stub_h2d_bytes()), "hash": "%x" % stub.stub_h2d_hash(),
                             "fronts": int(st["n_fronts"]), "nnz_L": int(st["nnz_L"])}


prepare(synth.graph_from_drive(synth.trackdrive(1)), "c1")
prepare(synth.graph_from_drive(synth.trackdrive(3, poses_per_lap=1500, seed=5)), "three_laps")   # >= 4096 blocks: pool path
prepare(small_graph(synth, 150), "small")
# the topology of the opt-in localiser repair: every landmark and all but the last W poses fixed --
# thousands of inactive edges, a handful of free blocks, landmark edges with only their pose end free
g1 = synth.graph_from_drive(synth.trackdrive(1))
out["window"] = {}
for W in (1, 3, 10, 50):
    ctx.graph_load(g1)
    for v in list(g1.lm_ids) + list(g1.pose_ids[:len(g1.pose_ids) - W]):
        ctx.graph_set_fixed(int(v), True)
    n = ctx.graph_prepare()
    st = ctx.graph_stats()
    out["window"][str(W)] = {"n": int(n), "blocks": int(st["n_blocks"]), "offdiag": int(st["n_offdiag_blocks"]), "fronts": int(st["n_fronts"])}
ctx.graph_load(g1)
for v in list(g1.lm_ids) + list(g1.pose_ids):
    ctx.graph_set_fixed(int(v), True)
out["window"]["all_fixed"] = {"prepare": int(ctx.graph_prepare()), "optimize": int(ctx.graph_optimize_rc(2)[0])}
# host side of the replica batch (value interleaving, upload / download staging) and of the sharded assembly
gs = small_graph(synth, 150)
pe, le, ez, oz = synth.perturb_replicas(gs, 8, seed=18)
ctx.graph_load(gs)
res = ctx.graph_optimize_batch(pe, le, oz, ez, iters=2)
out["batch"] = {"replicas": int(np.asarray(res[0]).shape[0])}
ctx.graph_load(g1)
ctx.graph_prepare_assembly_only()
P1 = len(g1.pose_ids)
out["shards"] = []
for p0, p1 in ((0, P1 // 2), (P1 // 2, P1)):
    l0, l1 = ctx.graph_shard_landmarks(p0, p1)
    ctx.graph_assemble_async(p0, p1)
    out["shards"].append([int(l0), int(l1)])
ctx.sync()
# incremental API, gauge flags, error paths (int status, message on the context, nothing thrown)
ctx.graph_clear()
z = np.zeros(3)
info = np.eye(3).ravel().copy()
rc = out["rc"]
rc["add_pose"] = L.slam_b200_graph_add_pose(ctx.h, 1000, 0.0, 0.0, 0.0)
rc["add_pose_duplicate"] = L.slam_b200_graph_add_pose(ctx.h, 1000, 0.0, 0.0, 0.0)
rc["add_landmark"] = L.slam_b200_graph_add_landmark(ctx.h, 0, 1.0, 1.0)
rc["edge_unknown_landmark"] = L.slam_b200_graph_add_edge_se2_xy(ctx.h, 1000, 7, z.ctypes.data_as(c_dp), info.ctypes.data_as(c_dp))
rc["edge_null_pointers"] = L.slam_b200_graph_add_edge_se2_xy(ctx.h, 1000, 0, None, None)
rc["set_fixed_unknown"] = L.slam_b200_graph_set_fixed(ctx.h, 4242, 1)
rc["last_error_set"] = int(bool(L.slam_b200_last_error(ctx.h)))
rc["prepare_no_edges"] = L.slam_b200_graph_prepare(ctx.h)
rc["optimize_nothing_to_do"] = ctx.graph_optimize_rc(3)[0]           # g2o: -1
# bulk load: a null array with a non-zero count is an argument error, and a load that fails half way (unknown id in
# the last edge, duplicate vertex id) leaves an EMPTY graph, not one a later optimise would silently run on
ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
dp = lambda a: a.ctypes.data_as(c_dp)
pid = np.array([1000, 1001, 1002], dtype=np.int32); pest = np.zeros(9)
lid = np.array([0, 1], dtype=np.int32); lest = np.ones(4)
eof = np.array([1000, 1001], dtype=np.int32); eot = np.array([1001, 1002], dtype=np.int32)
eoz = np.zeros(6); eoi = np.tile(np.eye(3).ravel(), 2)
elp = np.array([1000, 1001, 1002], dtype=np.int32); ell = np.array([0, 1, 1], dtype=np.int32)
elz = np.zeros(6); eli = np.tile(np.eye(2).ravel(), 3)
fx = np.array([1000], dtype=np.int32)


def load(pose_ids=pid, eo_z=eoz, el_lm=ell, fixed=fx, null_eo_info=False):
    return L.slam_b200_graph_load(ctx.h, 3, None if pose_ids is None else ip(pose_ids), dp(pest), 2, ip(lid), dp(lest),
                                  2, ip(eof), ip(eot), None if eo_z is None else dp(eo_z), None if null_eo_info else dp(eoi),
                                  3, ip(elp), ip(el_lm), dp(elz), dp(eli), 1, None if fixed is None else ip(fixed))


rc["load_ok"] = load()
nverts = lambda: L.slam_b200_graph_num_poses(ctx.h) + L.slam_b200_graph_num_landmarks(ctx.h)
rc["load_ok_vertices"] = nverts()
rc["load_null_pose_ids"] = load(pose_ids=None)
rc["load_null_eo_z"] = load(eo_z=None)
rc["load_null_eo_info"] = load(null_eo_info=True)
rc["load_null_fixed"] = load(fixed=None)
rc["load_unknown_landmark_in_last_edge"] = load(el_lm=np.array([0, 1, 9], dtype=np.int32))
rc["vertices_after_failed_load"] = nverts()
rc["optimize_after_failed_load"] = ctx.graph_optimize_rc(2)[0]        # empty graph: nothing to optimise, -1
rc["load_duplicate_id"] = load(pose_ids=np.array([1000, 1001, 1001], dtype=np.int32))
rc["vertices_after_duplicate"] = nverts()
rc["load_unknown_fixed_id"] = load(fixed=np.array([77], dtype=np.int32))
rc["vertices_after_unknown_fixed"] = nverts()
# map + frame staging paths (kernels are no-ops; buffers, growth and copies are real)
f = synth.cone_field(n_map=5000, n_obs=700, seed=4)
ctx.map_clear()
ctx.map_append(f.map_x, f.map_y, f.map_type)
rc["map_size"] = ctx.map_size()
x, y, t = ctx.map_read()
rc["map_roundtrip"] = int(np.array_equal(x, f.map_x) and np.array_equal(y, f.map_y) and np.array_equal(t, f.map_type))
fr = np.asfortranarray(np.random.default_rng(0).normal(size=(4, 900)))
ctx.cones_to_global(fr, np.zeros(3))
# kernels are no-ops under the stub: nothing would ever publish the frame's completion word, so the frame calls go
# through the copy path here (the staging buffers and status codes are what this case covers)
os.environ["SLAM_B200_FRAME_COPIES"] = "1"
ctx.assoc_localize_frame(fr, np.zeros(3), 1.2, 0)
ctx.close()
print(json.dumps(out))

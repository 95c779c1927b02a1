"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, and refuses to run without a GPU instead of falling back to the CPU."""
import ctypes
import os

import pytest

from conftest import skip_if_sanitizer_runtime_unusable


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.capi.lib()
    names = pkg.capi.exported_symbols()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.slam_b200_version() == 100


def test_no_cpu_fallback_without_a_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.SlamB200Error):
        pkg.Context(0)


def test_product_does_not_link_or_import_the_oracle(pkg):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pdir = os.path.join(root, "opendlv-logic-cfsd18-sensation-slam_b200")
    for dirpath, _, files in os.walk(pdir):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
                assert "orc_" not in txt, f


def test_host_code_over_a_stub_runtime(pkg, tmp_path):
    """The HOST code behind the C ABI (argument checks and status codes, graph_load, structure pass, symbolic
    phase, launch lists, packing of the structure upload, map / frame staging) on a box without a GPU: a
    subprocess with tests/stub_cudart.cpp LD_PRELOADed in place of the CUDA runtime (device memory = host heap,
    kernels = no-ops, nothing computed -- test infrastructure, the product never sees it).  The bytes uploaded for a
    graph must not depend on the size of the host pool; error paths return SLAM_B200_E_ARG with a message and
    g2o's -1 for "nothing to optimise"."""
    import json
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    pkg.build()
    stub = str(tmp_path / "stub_cudart.so")
    b = subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-o", stub, os.path.join(here, "stub_cudart.cpp")], capture_output=True, text=True)
    assert b.returncode == 0, b.stderr[-2000:]
    runs = []
    for threads in ("1", "8"):
        env = dict(os.environ, LD_PRELOAD=stub, SLAM_STUB_CUDART=stub, SLAM_B200_SYM_THREADS=threads)
        r = subprocess.run([sys.executable, os.path.join(here, "host_case_stub_runtime.py")], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-3000:]
        runs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert runs[0]["uploads"] == runs[1]["uploads"]                       # same structure bit for bit, 1 or 8 host threads
    up = runs[0]["uploads"]
    assert up["c1"]["n"] == 3590 and up["three_laps"]["n"] == 14090 and all(v["bytes"] > 0 for v in up.values())
    for W in (1, 3, 10, 50):                                              # the sliding window of the localiser repair
        w = runs[0]["window"][str(W)]
        assert w["n"] == 3 * W and w["blocks"] == W and w["offdiag"] == W - 1 and w["fronts"] >= 1, (W, w)
    assert runs[0]["window"]["all_fixed"] == {"prepare": 0, "optimize": -1}   # g2o: nothing to optimise
    assert runs[0]["batch"] == {"replicas": 8} and len(runs[0]["shards"]) == 2   # batch and sharded-assembly host paths ran
    assert all(0 <= l0 <= l1 for l0, l1 in runs[0]["shards"])
    rc = runs[0]["rc"]
    E_ARG = -101
    assert rc["add_pose"] == 0 and rc["add_landmark"] == 0 and rc["prepare_no_edges"] == 0
    assert rc["add_pose_duplicate"] == rc["edge_unknown_landmark"] == rc["edge_null_pointers"] == rc["set_fixed_unknown"] == E_ARG
    assert rc["last_error_set"] == 1 and rc["optimize_nothing_to_do"] == -1
    assert rc["map_size"] == 5000 and rc["map_roundtrip"] == 1
    # bulk load: null arrays -> E_ARG (no crash); a load failing half way leaves an empty graph
    assert rc["load_ok"] == 0 and rc["load_ok_vertices"] == 5
    assert rc["load_null_pose_ids"] == rc["load_null_eo_z"] == rc["load_null_eo_info"] == rc["load_null_fixed"] == E_ARG
    assert rc["load_unknown_landmark_in_last_edge"] == rc["load_duplicate_id"] == rc["load_unknown_fixed_id"] == E_ARG
    assert rc["vertices_after_failed_load"] == rc["vertices_after_duplicate"] == rc["vertices_after_unknown_fixed"] == 0
    assert rc["optimize_after_failed_load"] == -1


def test_host_code_under_sanitizers(tmp_path):
    """profiles/tools/host_sanitize.sh: the same host case on an AddressSanitizer + UBSan build of the library's
    host code (device code untouched), over the stub runtime whose "device" buffers are heap blocks ASan guards --
    an out-of-bounds structure array, a mis-sized staging buffer or a signed overflow in the index arithmetic of
    graph_load / the structure pass / the symbolic phase / the upload packing aborts with a report."""
    import json
    import shutil
    import subprocess
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["sh", os.path.join(root, "profiles", "tools", "host_sanitize.sh")], capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, OUT=str(tmp_path)))
    skip_if_sanitizer_runtime_unusable(r.stdout + r.stderr)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-4000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["uploads"]["three_laps"]["n"] == 14090 and out["rc"]["map_roundtrip"] == 1


def test_two_contexts_on_two_host_threads_under_tsan(tmp_path):
    """profiles/tools/host_sanitize.sh tsan: a ThreadSanitizer build of the library's host code over the stub runtime;
    two host threads, each with its own context and graph, run graph_load / graph_prepare / graph_optimize / map
    staging at the same time (tests/tsan_two_contexts_driver.cpp).  No data race, no hang, and each thread's analysis
    equals what it gets alone -- the threading contract of include/slam_b200.h on everything but the kernels (the
    GPU half is test_graph_gpu.py::test_two_contexts_on_two_host_threads)."""
    import shutil
    import subprocess
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["sh", os.path.join(root, "profiles", "tools", "host_sanitize.sh"), "tsan"], capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, OUT=str(tmp_path)))
    skip_if_sanitizer_runtime_unusable(r.stdout + r.stderr)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1].startswith("ok "), r.stdout[-1500:] + r.stderr[-4000:]

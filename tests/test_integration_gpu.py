"""The drop-in boundary inside the reference tree, on the B200.

integration/_build/patched_public_replay = the reference's own src/slam.hpp + slam.cpp with integration/slam_b200.patch
applied (back half of class Slam -> C ABI), its cone.cpp, cluon and the generated message set, linked against
libslam_b200.so; built in the build container by integration/build.sh (the reference tree does not exist on the GPU
box, the binary travels like the built .so).  It is driven through the PUBLIC interface only -- OD4 session constructor,
nextPose / nextYawRate / nextCone with cluon Envelopes, draw* -- and must answer what the UNMODIFIED reference answered
on the same Envelopes (tests/golden/public_api_reference.npz): association rows (drawGraph) and dropped frames
identical, stored poses bit for bit, optimised map within 1e-6 relative, NaN cones in the same slots."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
EXE = os.path.join(ROOT, "integration", "_build", "patched_public_replay")

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["odd", "c1", "loop1", "loop2", "loop5", "nan0", "nan2", "yaw1"])
def test_patched_reference_tree_equals_the_reference_through_the_public_api(pkg, name):
    import make_public_api_golden as mk
    if not os.path.exists(EXE):
        pytest.fail("integration/_build/patched_public_replay is missing: run integration/build.sh where /root/reference exists")
    g = np.load(os.path.join(GOLD, "public_api_reference.npz"))
    r = mk.run(EXE, os.path.join(GOLD, "public_api_drive_%s.bin" % name))
    assert np.array_equal(r["frames"], g[name + "_frames"])            # the same frames stored / dropped (200 m gate)
    assert int(r["dropped"][0]) == int(g[name + "_dropped"][0])
    assert np.array_equal(r["graph_ptr"], g[name + "_graph_ptr"])
    assert np.array_equal(r["graph_ids"], g[name + "_graph_ids"])      # every association decision, in order
    assert np.array_equal(r["poses"], g[name + "_poses"])              # incl. the heading correction: bit for bit
    m, gm = r["map"], g[name + "_map"]
    assert m.shape == gm.shape
    assert np.array_equal(m[:, 2:], gm[:, 2:])                         # type, id
    assert np.array_equal(np.isnan(m[:, 0]), np.isnan(gm[:, 0]))       # NaN cones (absent objectIds) in the same slots
    ok = ~np.isnan(gm[:, 0])
    scale = max(1.0, float(np.max(np.abs(gm[ok, :2]))))
    assert np.max(np.abs(m[ok, :2] - gm[ok, :2])) <= 1e-6 * scale      # optimised landmarks (north_star: 1e-6 relative)
    assert np.max(np.abs(r["current"] - g[name + "_current"])) <= 1e-6 * scale   # drawCurrentPose(): the sent pose

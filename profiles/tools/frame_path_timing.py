"""Per-frame path of the drop-in: the C1 drive through the raw C-ABI frame call and through the
drop-in Slam class, beside the reference's real slam.cpp on a host core.  SLAM_B200_FRAME_COPIES=1 = the first version
(H2D copy + kernel + two D2H copies + stream synchronisation per frame) instead of the mapped mailbox."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
import bench  # noqa: E402

pkg = bench.load_pkg() if hasattr(bench, "load_pkg") else __import__("__graft_entry__").load_package()
pkg.build()
out = bench.bench_frame_assoc(pkg, torch, 0, pkg.synth.trackdrive(1))
out["frame_copies"] = bool(os.environ.get("SLAM_B200_FRAME_COPIES"))
print(json.dumps(out))

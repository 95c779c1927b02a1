"""Times the bulk grid association kernel (config 4) with the L2 flushed between launches and with
the map left hot in L2, to separate DRAM random-access cost from the dependent-chain latency."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
f = pkg.synth.cone_field()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = pkg.Context(0, stream=stream.cuda_stream)
ctx.map_append(f.map_x, f.map_y, f.map_type)
ctx.map_build_grid(1.2)
n = f.frame.shape[1]
d_in = torch.from_numpy(np.ascontiguousarray(f.frame.T)).to(dev)
d_out = torch.empty(n, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(do_flush, reps=20):
    ts = []
    with torch.cuda.stream(stream):
        for _ in range(reps):
            if do_flush:
                flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.assoc_bulk_dev(d_in.data_ptr(), n, f.pose, 1.2, 0, 1, d_out.data_ptr())
            e1.record(stream)
            ts.append((e0, e1))
    stream.synchronize()
    v = sorted(a.elapsed_time(b) * 1e3 for a, b in ts)
    return v[len(v) // 2], v[0]


for _ in range(3):
    run(False, 3)
print("L2 flushed : median %.2f us, min %.2f us" % run(True))
print("L2 hot     : median %.2f us, min %.2f us" % run(False))
# sorted observations (spatially coherent order, like a real lidar sweep): same work, better locality
g, _ = ctx.cones_to_global(f.frame, f.pose)
order = np.lexsort((g[:, 0], np.floor(g[:, 1] / 2.4)))
d_in2 = torch.from_numpy(np.ascontiguousarray(f.frame.T[order])).to(dev)
d_in, d_keep = d_in2, d_in
print("row-sorted obs, L2 flushed : median %.2f us, min %.2f us" % run(True))

"""Grid association kernel duration vs number of observations (map of 1M cones, L2 flushed):
separates the fixed launch/dependent-chain latency from the per-observation cost."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
f = pkg.synth.cone_field(n_obs=800_000)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
ctx = pkg.Context(0, stream=stream.cuda_stream)
ctx.map_append(f.map_x, f.map_y, f.map_type)
ctx.map_build_grid(1.2)
d_in = torch.from_numpy(np.ascontiguousarray(f.frame.T)).to(dev)
d_out = torch.empty(f.frame.shape[1], dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (1, 1000, 10_000, 50_000, 100_000, 200_000, 400_000, 800_000):
    ts = []
    with torch.cuda.stream(stream):
        for _ in range(12):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.assoc_bulk_dev(d_in.data_ptr(), n, f.pose, 1.2, 0, 1, d_out.data_ptr())
            e1.record(stream)
            ts.append((e0, e1))
    stream.synchronize()
    v = sorted(a.elapsed_time(b) * 1e3 for a, b in ts)[2:]
    bytes_alg = 36.0 * n + 20.0 * 1_000_000
    print("n=%7d  median %7.2f us  -> %6.2f ns/obs, %5.1f%% of HBM roofline (36n+20M bytes)"
          % (n, v[len(v) // 2], v[len(v) // 2] * 1e3 / n, 100 * bytes_alg / (v[len(v) // 2] * 1e-6) / 6550.7e9))

"""Body of test_graph_gpu.py::test_two_contexts_on_two_host_threads, run as its own process: the C ABI allows any
number of contexts, each used by one host thread at a time (include/slam_b200.h).  Two threads, each with its own
context and its own graph, load / analyse / optimise at the same time (ctypes releases the GIL during the calls);
each must get bit for bit what it gets alone.  Exit code 0 = equal."""
import os
import sys
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from conftest import load_pkg, small_graph  # noqa: E402


def main():
    pkg = load_pkg()
    pkg.build()
    synth = pkg.synth
    graphs = [synth.graph_from_drive(synth.trackdrive(5, poses_per_lap=1000, seed=31)),   # 5,000 poses: pool path of the analysis
              small_graph(synth, 400, seed=9)]

    def solve(g, rounds):
        ctx = pkg.Context(0)
        out = None
        for _ in range(rounds):
            ctx.graph_load(g)                    # new topology version: structure pass + symbolic phase + upload
            n, chi2 = ctx.graph_optimize(10)
            pe, le = ctx.graph_get_estimates()
            out = (n, chi2.copy(), pe.copy(), le.copy())
        ctx.close()
        return out

    alone = [solve(g, 1) for g in graphs]
    together = [None, None]
    errors = []

    def worker(k):
        try:
            together[k] = solve(graphs[k], 6 if k == 0 else 40)
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    for k in range(2):
        a, b = alone[k], together[k]
        assert a[0] == b[0] == 10, (k, a[0], b[0])
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), k
    print("two contexts on two host threads: results identical to the single-threaded runs")


if __name__ == "__main__":
    main()

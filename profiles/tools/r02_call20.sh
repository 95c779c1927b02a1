#!/bin/bash
# C3: register-cap levels of the register-resident tile kernels + the 8-tile-row backward kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs -k "batched or replica or tile" > gpurun_out/r02_call20_tests.log 2>&1
for t in 0 1 2 3; do
SLAM_B200_TILE_TIGHT=$t python bench.py --workload c3 --steps 5 > gpurun_out/r02_call20_c3_tight$t.json 2> gpurun_out/r02_call20_c3_tight$t.err
done

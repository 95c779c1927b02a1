#!/bin/bash
# C3: register caps of the register-resident tile kernels (0 = none, 1 = T <= 5 at 20 warps, 2 = all tight)
set -x
mkdir -p gpurun_out
for t in 0 1 2; do
SLAM_B200_TILE_TIGHT=$t python bench.py --workload c3 --steps 5 > gpurun_out/r02_call19_c3_tight$t.json 2> gpurun_out/r02_call19_c3_tight$t.err
done

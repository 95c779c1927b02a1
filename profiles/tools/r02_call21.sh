#!/bin/bash
set -x
mkdir -p gpurun_out
for t in 2111 1211 1121 2211 0111 1011 1101 2011; do
SLAM_B200_TILE_TIGHT=$t python bench.py --workload c3 --steps 5 > gpurun_out/r02_call21_c3_t$t.json 2> gpurun_out/r02_call21_c3_t$t.err
done

#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:factor_tile_reg_kernel --launch-skip 0 --launch-count 3 -f -o gpurun_out/r02_prof_factor_tile_c3 python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/ncu_c3f.log 2>&1

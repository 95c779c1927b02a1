#!/bin/bash
# round 2, call 4: ncu of the tiled kernels on config 3 (2,048 replicas keep the capture short)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_c3_launches.csv python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/ncu_c3l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:factor_tile_kernel --launch-skip 80 --launch-count 3 -f -o gpurun_out/r02_prof_factor_tile_c3 python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/ncu_c3f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:backward_tile_kernel --launch-skip 117 --launch-count 3 -f -o gpurun_out/r02_prof_backward_tile_c3 python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/ncu_c3b.log 2>&1
ls -la gpurun_out/*.ncu-rep

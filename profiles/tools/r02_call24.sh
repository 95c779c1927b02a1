#!/bin/bash
# round-2 ncu evidence: launch lists (C2 headline, C3) + full captures of the kernels that changed
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_c2_launches.csv python bench.py --no-assoc --no-sharded --steps 2 --warmup 3 > gpurun_out/r02_call24_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_c3_launches.csv python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/r02_call24_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:factor_tile_reg_tight_kernel --launch-skip 0 --launch-count 3 -f -o gpurun_out/r02_prof_factor_tile_tight_c3 python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/r02_call24_ncu_c3f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:backward_tile_kernel --launch-skip 15 --launch-count 1 -f -o gpurun_out/r02_prof_backward_tile_c3 python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/r02_call24_ncu_c3b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:drive_replicas_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_prof_drive_replicas python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/r02_call24_ncu_drv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assoc_bulk_grid_batched_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_prof_assoc_batched_c4 python bench.py --workload c4 --steps 24 > gpurun_out/r02_call24_ncu_c4.log 2>&1

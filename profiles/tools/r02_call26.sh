#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c2_launches_f3.csv python bench.py --no-assoc --no-sharded --steps 2 --warmup 3 > gpurun_out/r02_call26_f3.log 2>&1
SLAM_B200_FACTOR_VARIANT=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c2_launches_f2.csv python bench.py --no-assoc --no-sharded --steps 2 --warmup 3 > gpurun_out/r02_call26_f2.log 2>&1

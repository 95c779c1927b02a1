#!/bin/bash
# launch list of config 3 (2,048 replicas): every launch of the first GN iterations with its duration
set -x
mkdir -p gpurun_out
python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/r02_call9_c3.json 2> gpurun_out/r02_call9_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_c3_launches.csv python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/r02_call9_ncu.log 2>&1

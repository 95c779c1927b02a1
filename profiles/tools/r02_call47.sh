#!/bin/bash
# round 2, call 47: ncu evidence for the final single-graph kernels -- launch list of the C2 command and a full capture of
# the ten factor2_kernel launches of one Gauss-Newton iteration (each only after the plain run exited 0)
set -x
mkdir -p gpurun_out
python bench.py --no-assoc --no-sharded --steps 2 --warmup 3 > gpurun_out/r02_call47_plain.json 2> gpurun_out/r02_call47_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_c2_launches_final.csv python bench.py --no-assoc --no-sharded --steps 2 --warmup 3 > gpurun_out/r02_call47_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:factor2_kernel --launch-skip 40 --launch-count 10 -f -o gpurun_out/r02_prof_factor2_c2 python bench.py --no-assoc --no-sharded --steps 2 --warmup 3 > gpurun_out/r02_call47_ncu_full.log 2>&1
ls -la gpurun_out/r02_prof_factor2_c2.ncu-rep

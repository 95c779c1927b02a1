#!/bin/bash
# round 2, call 53: why is the cold call slower inside bench.py than in e2e_breakdown.py on the same box? (library thread pools)
set -x
mkdir -p gpurun_out
python profiles/tools/e2e_breakdown.py > gpurun_out/r02_call53_e2e.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call53_c2_default.json 2> gpurun_out/r02_call53_c2_default.err
OMP_NUM_THREADS=1 MKL_NUM_THREADS=1 OPENBLAS_NUM_THREADS=1 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call53_c2_omp1.json 2> gpurun_out/r02_call53_c2_omp1.err
KMP_BLOCKTIME=0 GOMP_SPINCOUNT=0 OMP_WAIT_POLICY=passive python bench.py --no-assoc --no-sharded > gpurun_out/r02_call53_c2_passive.json 2> gpurun_out/r02_call53_c2_passive.err
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call53_c2_default2.json 2> gpurun_out/r02_call53_c2_default2.err

#!/bin/bash
# factor3_kernel (register-resident 16-warp front kernel) against factor2_kernel on the C2 headline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call25_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call25_c2_f3.json 2> gpurun_out/r02_call25_c2_f3.err
SLAM_B200_FACTOR_VARIANT=2 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call25_c2_f2.json 2> gpurun_out/r02_call25_c2_f2.err

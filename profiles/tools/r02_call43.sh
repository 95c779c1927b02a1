#!/bin/bash
# round 2, call 43: trailing update of factor2_kernel on the fp64 tensor pipe (8 x 8 tiles from shared memory), padded stride
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call43_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call43_c2.json 2> gpurun_out/r02_call43_c2.err
SLAM_B200_UPDATE_MMA=0 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call43_c2_scalar.json 2> gpurun_out/r02_call43_c2_scalar.err
python profiles/tools/front_timeline.py > gpurun_out/r02_call43_timeline.log 2>&1
SLAM_B200_DBG_FRONT=2218 SLAM_B200_NO_PDL=1 SLAM_B200_PHASE_CLOCKS=1 SLAM_B200_NO_CUDA_GRAPH=1 python profiles/tools/factor_phase_clocks.py > gpurun_out/r02_call43_clocks_f2218.log 2>&1

#!/bin/bash
# round 2, call 34: L1 prefetch of the static lists before the wait + flat batched extend-add in factor2_kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call34_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call34_c2.json 2> gpurun_out/r02_call34_c2.err
SLAM_B200_PHASE_CLOCKS=1 SLAM_B200_NO_CUDA_GRAPH=1 python profiles/tools/factor_phase_clocks.py > gpurun_out/r02_call34_f2_clocks.log 2>&1

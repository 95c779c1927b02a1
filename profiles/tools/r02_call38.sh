#!/bin/bash
# round 2, call 38: tests after the merged-level fix; phase clocks of the level-8 merge front (s=42, u=92) with and
# without warp 0's scheduler mates in the update
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call38_tests.log 2>&1
for la in 0 1000; do
SLAM_B200_LA_IDLE=$la SLAM_B200_DBG_FRONT=2218 SLAM_B200_NO_PDL=1 SLAM_B200_PHASE_CLOCKS=1 SLAM_B200_NO_CUDA_GRAPH=1 python profiles/tools/factor_phase_clocks.py > gpurun_out/r02_call38_clocks_f2218_la$la.log 2>&1
SLAM_B200_LA_IDLE=$la SLAM_B200_DBG_FRONT=2219 SLAM_B200_NO_PDL=1 SLAM_B200_PHASE_CLOCKS=1 SLAM_B200_NO_CUDA_GRAPH=1 python profiles/tools/factor_phase_clocks.py > gpurun_out/r02_call38_clocks_f2219_la$la.log 2>&1
done

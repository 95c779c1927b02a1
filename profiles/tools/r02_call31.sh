#!/bin/bash
set -x
mkdir -p gpurun_out
SLAM_B200_FACTOR_VARIANT=3 SLAM_B200_PHASE_CLOCKS=1 SLAM_B200_NO_CUDA_GRAPH=1 python profiles/tools/factor_phase_clocks.py > gpurun_out/r02_call31_f3_clocks.log 2>&1

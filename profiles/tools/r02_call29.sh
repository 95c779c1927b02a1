#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call29_tests.log 2>&1
python profiles/tools/e2e_breakdown.py > gpurun_out/r02_call29_e2e.log 2>&1
SLAM_B200_SYM_DEBUG=1 python profiles/tools/e2e_breakdown.py 2>&1 | grep -v "region [0-9]" | tail -40 > gpurun_out/r02_call29_e2e_debug.log

#!/bin/bash
# round 2, call 48: the new variant / timeline tests, cold-call breakdown on the final build
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs -k "variants or trailing_update or timeline" > gpurun_out/r02_call48_tests.log 2>&1
python profiles/tools/e2e_breakdown.py > gpurun_out/r02_call48_e2e.log 2>&1
SLAM_B200_SYM_DEBUG=1 python profiles/tools/e2e_breakdown.py > gpurun_out/r02_call48_e2e_debug.log 2>&1
nproc > gpurun_out/r02_call48_nproc.log; lscpu | head -20 >> gpurun_out/r02_call48_nproc.log

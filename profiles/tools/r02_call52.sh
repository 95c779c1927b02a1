#!/bin/bash
# round 2, call 52: launch lists built beside the last stage of the symbolic analysis -- tests + cold-call breakdown
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call52_tests.log 2>&1
python profiles/tools/e2e_breakdown.py > gpurun_out/r02_call52_e2e.log 2>&1
SLAM_B200_SYM_DEBUG=1 python profiles/tools/e2e_breakdown.py 2>&1 | tail -40 > gpurun_out/r02_call52_e2e_debug.log
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call52_c2.json 2> gpurun_out/r02_call52_c2.err

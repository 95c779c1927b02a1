#!/bin/bash
# round 2, call 39: row maps + pivot rhs staged before the wait, next child's entries requested before the current one is added
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call39_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call39_c2.json 2> gpurun_out/r02_call39_c2.err
python profiles/tools/front_timeline.py > gpurun_out/r02_call39_timeline.log 2>&1

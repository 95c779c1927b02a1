#!/bin/bash
# round 2, call 58: warp-per-front backward kernel with the next step's L columns requested one step ahead
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call58_tests.log 2>&1
python bench.py --no-assoc > gpurun_out/r02_call58_default.json 2> gpurun_out/r02_call58_default.err
python profiles/tools/front_timeline.py > gpurun_out/r02_call58_timeline.log 2>&1

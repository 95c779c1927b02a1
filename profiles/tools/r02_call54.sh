#!/bin/bash
# round 2, call 54: extend-add batches in three shapes (2x1, 4x2, 6x3)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call54_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call54_c2.json 2> gpurun_out/r02_call54_c2.err
python profiles/tools/front_timeline.py > gpurun_out/r02_call54_timeline.log 2>&1
python profiles/tools/front_timeline.py 1 > gpurun_out/r02_call54_timeline_c1.log 2>&1

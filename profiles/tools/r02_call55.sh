#!/bin/bash
# round 2, call 55: extend-add with two alternating batches (requests issued two children ahead)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call55_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call55_c2.json 2> gpurun_out/r02_call55_c2.err
SLAM_B200_EA_SINGLE=1 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call55_c2_single.json 2> gpurun_out/r02_call55_c2_single.err
python profiles/tools/front_timeline.py > gpurun_out/r02_call55_timeline.log 2>&1
python profiles/tools/front_timeline.py 1 > gpurun_out/r02_call55_timeline_c1.log 2>&1

#!/bin/bash
# round 2, call 37: mixed levels as one launch, largest fronts first; phase timeline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call37_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call37_c2.json 2> gpurun_out/r02_call37_c2.err
SLAM_B200_NO_LEVEL_MERGE=1 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call37_c2_nomerge.json 2> gpurun_out/r02_call37_c2_nomerge.err
python profiles/tools/front_timeline.py > gpurun_out/r02_call37_timeline.log 2>&1
python profiles/tools/front_timeline.py 1 > gpurun_out/r02_call37_timeline_c1.log 2>&1

#!/bin/bash
# round 2, call 67: the first factor launch resident under the landmark assembly, the update behind the last backward launch (PDL)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py tests/test_peer_exchange_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call67_tests.log 2>&1
python bench.py --no-assoc > gpurun_out/r02_call67_default.json 2> gpurun_out/r02_call67_default.err
python profiles/tools/front_timeline.py > gpurun_out/r02_call67_timeline.log 2>&1

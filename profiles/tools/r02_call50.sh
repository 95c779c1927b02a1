#!/bin/bash
# round 2, call 50: mixed levels with more fronts than SMs as 256-thread CTAs (two per SM) instead of two waves of 512
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call50_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call50_c2.json 2> gpurun_out/r02_call50_c2.err
SLAM_B200_MERGE_THREADS=512 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call50_c2_512.json 2> gpurun_out/r02_call50_c2_512.err
python profiles/tools/front_timeline.py > gpurun_out/r02_call50_timeline.log 2>&1

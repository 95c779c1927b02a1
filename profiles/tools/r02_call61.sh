#!/bin/bash
# round 2, call 61: storage offsets of the write-out prefetched in the prologue
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call61_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call61_c2.json 2> gpurun_out/r02_call61_c2.err
python profiles/tools/front_timeline.py > gpurun_out/r02_call61_timeline.log 2>&1

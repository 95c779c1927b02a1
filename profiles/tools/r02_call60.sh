#!/bin/bash
# round 2, call 60: is the CUDA graph still worth its capture + instantiate on a cold call now that the levels are chained by PDL?
set -x
mkdir -p gpurun_out
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call60_graph.json 2> gpurun_out/r02_call60_graph.err
SLAM_B200_NO_CUDA_GRAPH=1 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call60_nograph.json 2> gpurun_out/r02_call60_nograph.err
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call60_graph2.json 2> gpurun_out/r02_call60_graph2.err
SLAM_B200_NO_CUDA_GRAPH=1 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call60_nograph2.json 2> gpurun_out/r02_call60_nograph2.err

#!/bin/bash
# round 2, call 63: thin levels of a replica batch as ONE multi-class launch, chained by programmatic dependent launch
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call63_tests.log 2>&1
python bench.py --workload c3 --steps 5 --replicas 512 > gpurun_out/r02_call63_c3_R512.json 2> gpurun_out/r02_call63_c3_R512.err
SLAM_B200_TILE_NO_MULTI=1 python bench.py --workload c3 --steps 5 --replicas 512 > gpurun_out/r02_call63_c3_R512_nomulti.json 2> gpurun_out/r02_call63_c3_R512_nomulti.err
python bench.py --workload c3 --steps 5 > gpurun_out/r02_call63_c3.json 2> gpurun_out/r02_call63_c3.err
SLAM_B200_TILE_NO_MULTI=1 python bench.py --workload c3 --steps 5 > gpurun_out/r02_call63_c3_nomulti.json 2> gpurun_out/r02_call63_c3_nomulti.err
python bench.py --workload c3 --steps 5 --replicas 1024 > gpurun_out/r02_call63_c3_R1024.json 2> gpurun_out/r02_call63_c3_R1024.err
SLAM_B200_TILE_NO_MULTI=1 python bench.py --workload c3 --steps 5 --replicas 1024 > gpurun_out/r02_call63_c3_R1024_nomulti.json 2> gpurun_out/r02_call63_c3_R1024_nomulti.err

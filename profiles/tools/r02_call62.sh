#!/bin/bash
# round 2, call 62: programmatic dependent launch along the backward sweep of the replica batches (backward_tile_kernel)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call62_tests.log 2>&1
python bench.py --workload c3 --steps 5 > gpurun_out/r02_call62_c3.json 2> gpurun_out/r02_call62_c3.err
python bench.py --workload c3 --steps 5 --replicas 512 > gpurun_out/r02_call62_c3_R512.json 2> gpurun_out/r02_call62_c3_R512.err
SLAM_B200_NO_PDL=1 python bench.py --workload c3 --steps 5 --replicas 512 > gpurun_out/r02_call62_c3_R512_nopdl.json 2> gpurun_out/r02_call62_c3_R512_nopdl.err
SLAM_B200_NO_PDL=1 python bench.py --workload c3 --steps 5 > gpurun_out/r02_call62_c3_nopdl.json 2> gpurun_out/r02_call62_c3_nopdl.err

#!/bin/bash
# fork/join of size classes per level + nested-dissection leaf size of the batch ordering; 4,096 and 512 replicas per GPU
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call10_tests.log 2>&1
for R in 4096 512; do
python bench.py --workload c3 --steps 5 --replicas $R > gpurun_out/r02_call10_c3_fork_R$R.json 2> gpurun_out/r02_call10_c3_fork_R$R.err
SLAM_B200_TILE_NO_FORK=1 python bench.py --workload c3 --steps 5 --replicas $R > gpurun_out/r02_call10_c3_nofork_R$R.json 2> gpurun_out/r02_call10_c3_nofork_R$R.err
for leaf in 64 128 256; do
SLAM_B200_ND_LEAF=$leaf python bench.py --workload c3 --steps 5 --replicas $R > gpurun_out/r02_call10_c3_leaf${leaf}_R$R.json 2> gpurun_out/r02_call10_c3_leaf${leaf}_R$R.err
done; done

#!/bin/bash
# round 2, call 33: programmatic dependent launch between the levels of the single-graph front kernels (parity tests,
# then C2 with / without it) and the nested-dissection region size swept on C2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_slam_host_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call33_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call33_c2_pdl.json 2> gpurun_out/r02_call33_c2_pdl.err
SLAM_B200_NO_PDL=1 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call33_c2_nopdl.json 2> gpurun_out/r02_call33_c2_nopdl.err
for leaf in 700 1500 2100 2600; do
  SLAM_B200_ND_LEAF=$leaf python bench.py --no-assoc --no-sharded > gpurun_out/r02_call33_c2_leaf$leaf.json 2> gpurun_out/r02_call33_c2_leaf$leaf.err
done

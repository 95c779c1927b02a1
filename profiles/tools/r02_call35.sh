#!/bin/bash
# round 2, call 35: column-batched extend-add; warp 0's scheduler mates idle below a row threshold (sweep)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call35_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call35_c2.json 2> gpurun_out/r02_call35_c2.err
for la in 48 72 100 1000; do
  SLAM_B200_LA_IDLE=$la python bench.py --no-assoc --no-sharded > gpurun_out/r02_call35_c2_la$la.json 2> gpurun_out/r02_call35_c2_la$la.err
done
SLAM_B200_PHASE_CLOCKS=1 SLAM_B200_NO_CUDA_GRAPH=1 python profiles/tools/factor_phase_clocks.py > gpurun_out/r02_call35_f2_clocks.log 2>&1

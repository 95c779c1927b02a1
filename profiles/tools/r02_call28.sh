#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call28_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call28_c2_f3.json 2> gpurun_out/r02_call28_c2_f3.err
SLAM_B200_FACTOR_VARIANT=2 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call28_c2_f2.json 2> gpurun_out/r02_call28_c2_f2.err
python bench.py --workload c3 --steps 5 > gpurun_out/r02_call28_c3.json 2> gpurun_out/r02_call28_c3.err
SLAM_B200_PHASE_CLOCKS=1 SLAM_B200_NO_CUDA_GRAPH=1 python profiles/tools/factor_phase_clocks.py > gpurun_out/r02_call28_f3_clocks.log 2>&1

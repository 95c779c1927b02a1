#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call8_tests.log 2>&1
python bench.py --workload c3 --steps 5 > gpurun_out/r02_call8_c3_reg.json 2> gpurun_out/r02_call8_c3_reg.err
SLAM_B200_TILE_NO_REG=1 python bench.py --workload c3 --steps 5 > gpurun_out/r02_call8_c3_smem.json 2> gpurun_out/r02_call8_c3_smem.err
SLAM_B200_TILE_WPC_F=2 python bench.py --workload c3 --steps 5 > gpurun_out/r02_call8_c3_reg_w2.json 2> gpurun_out/r02_call8_c3_reg_w2.err

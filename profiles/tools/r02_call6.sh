#!/bin/bash
# tiled kernels: parity tests + config 3 with 1 / 2 / 4 warps per CTA
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call6_tests.log 2>&1
for wf in 1 2 4; do for wb in 1 4; do
SLAM_B200_TILE_WPC_F=$wf SLAM_B200_TILE_WPC_B=$wb python bench.py --workload c3 --steps 5 > gpurun_out/r02_call6_c3_f${wf}_b${wb}.json 2> gpurun_out/r02_call6_c3_f${wf}_b${wb}.err
done; done

#!/bin/bash
# round 2, call 3: tiled batched factorisation -- parity tests, then config 3 (tile path vs the warp-per-front path)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call3_tests.log 2>&1
python bench.py --workload c3 --steps 5 > gpurun_out/r02_call3_c3_tile.json 2> gpurun_out/r02_call3_c3_tile.err
SLAM_B200_NO_TILE_PATH=1 python bench.py --workload c3 --steps 5 > gpurun_out/r02_call3_c3_old.json 2> gpurun_out/r02_call3_c3_old.err

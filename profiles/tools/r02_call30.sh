#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call30_tests.log 2>&1
python bench.py --workload c3 --steps 5 > gpurun_out/r02_call30_c3.json 2> gpurun_out/r02_call30_c3.err
python bench.py --workload c3 --steps 5 --replicas 512 > gpurun_out/r02_call30_c3_R512.json 2> gpurun_out/r02_call30_c3_R512.err

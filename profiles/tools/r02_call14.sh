#!/bin/bash
# config 5 solve (1 GPU): corridor parity test + the c5 bench line with its solve leg
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs -k "corridor or c2_full" > gpurun_out/r02_call14_tests.log 2>&1
timeout 1200 python bench.py --workload c5 --steps 10 > gpurun_out/r02_call14_c5.json 2> gpurun_out/r02_call14_c5.err

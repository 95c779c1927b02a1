#!/bin/bash
set -x
mkdir -p gpurun_out
python profiles/tools/corridor_solve_check.py 30000 > gpurun_out/r02_call15_corridor30k.json 2> gpurun_out/r02_call15_corridor30k.err
python profiles/tools/corridor_solve_check.py 3000 > gpurun_out/r02_call15_corridor3k.json 2> gpurun_out/r02_call15_corridor3k.err
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q -x -rxXs -k "corridor" > gpurun_out/r02_call15_tests.log 2>&1

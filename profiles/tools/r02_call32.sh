#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_assoc_gpu.py tests/test_integration_gpu.py -m gpu -q -x -rxXs > gpurun_out/r02_call32_tests.log 2>&1
timeout 900 python bench.py --workload c4 --steps 24 > gpurun_out/r02_call32_c4.json 2> gpurun_out/r02_call32_c4.err

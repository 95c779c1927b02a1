#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rxXs > gpurun_out/r02_call16_tests.log 2>&1
timeout 1200 python bench.py --workload c5 --steps 10 > gpurun_out/r02_call16_c5.json 2> gpurun_out/r02_call16_c5.err
python profiles/tools/frame_path_timing.py > gpurun_out/r02_call16_frame.json 2> gpurun_out/r02_call16_frame.err

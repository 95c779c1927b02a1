#!/bin/bash
# round 2, call 59: final build -- whole GPU suite, smoke, default line + reference arm as the driver runs them
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rxXs > gpurun_out/r02_call59_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_call59_smoke.log 2>&1
( time python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_bench_default_1gpu_final.json 2> gpurun_out/r02_bench_default_1gpu_final.err ) 2> gpurun_out/r02_bench_default_1gpu_final.time
( time python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_1gpu_final.json 2> gpurun_out/r02_bench_reference_1gpu_final.err ) 2> gpurun_out/r02_bench_reference_1gpu_final.time
python profiles/tools/front_timeline.py > gpurun_out/r02_call59_timeline.log 2>&1
python profiles/tools/front_timeline.py 1 > gpurun_out/r02_call59_timeline_c1.log 2>&1

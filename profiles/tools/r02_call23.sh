#!/bin/bash
# whole GPU suite + the default bench line (what the driver runs) + the reference arm, timed
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rxXs > gpurun_out/r02_call23_tests.log 2>&1
( time python bench.py > gpurun_out/r02_call23_bench_default.json 2> gpurun_out/r02_call23_bench_default.err ) 2> gpurun_out/r02_call23_bench_default.time
( time python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_call23_bench_reference.json 2> gpurun_out/r02_call23_bench_reference.err ) 2> gpurun_out/r02_call23_bench_reference.time
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_call23_smoke.log 2>&1

#!/bin/bash
# round 2, call 46: the default line (all sections) on the new single-graph kernels + timelines
set -x
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r02_call46_default.json 2> gpurun_out/r02_call46_default.err ) 2> gpurun_out/r02_call46_default.time
python profiles/tools/front_timeline.py > gpurun_out/r02_call46_timeline.log 2>&1
python profiles/tools/front_timeline.py 1 > gpurun_out/r02_call46_timeline_c1.log 2>&1

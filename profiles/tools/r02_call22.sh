#!/bin/bash
# cold call: ND with one sweep less; how much of the host time is allocator churn (glibc thresholds raised by environment)
set -x
mkdir -p gpurun_out
python profiles/tools/e2e_breakdown.py > gpurun_out/r02_call22_e2e.log 2>&1
MALLOC_MMAP_THRESHOLD_=1073741824 MALLOC_TRIM_THRESHOLD_=2147483647 MALLOC_TOP_PAD_=268435456 python profiles/tools/e2e_breakdown.py > gpurun_out/r02_call22_e2e_malloc.log 2>&1
SLAM_B200_SYM_DEBUG=1 MALLOC_MMAP_THRESHOLD_=1073741824 MALLOC_TRIM_THRESHOLD_=2147483647 MALLOC_TOP_PAD_=268435456 python profiles/tools/e2e_breakdown.py 2>&1 | grep -v "region [0-9]" | tail -45 > gpurun_out/r02_call22_e2e_malloc_debug.log

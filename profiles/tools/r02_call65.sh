#!/bin/bash
# round 2, call 65: host pool size against the cold call (16 hardware threads on the box)
set -x
mkdir -p gpurun_out
for t in 16 12 8 6 4; do
  SLAM_B200_SYM_THREADS=$t python profiles/tools/e2e_breakdown.py > gpurun_out/r02_call65_e2e_t$t.log 2>&1
done

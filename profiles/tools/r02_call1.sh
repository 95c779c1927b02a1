#!/bin/bash
# round 2, call 1: sanitizer logs (memcheck / racecheck / synccheck), the -m gpu suite with xfail detail,
# and wall-clock of the sharded workloads (setup cost decides whether they can ride in the default bench line)
set -x
mkdir -p gpurun_out
export SLAM_B200_SANITIZER_SMALL=1
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python profiles/tools/sanitizer_case.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r02_sanitizer_$tool.log
done
unset SLAM_B200_SANITIZER_SMALL
timeout 1500 python -m pytest tests -m gpu -q -rxXs > gpurun_out/r02_gputests_call1.log 2>&1
( time python bench.py --workload c3 --steps 5 ) > gpurun_out/r02_c3_call1.json 2> gpurun_out/r02_c3_call1.err
( time python bench.py --workload c5 --steps 10 ) > gpurun_out/r02_c5_call1.json 2> gpurun_out/r02_c5_call1.err
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02_c2_call1.json 2> gpurun_out/r02_c2_call1.err
nvidia-smi > gpurun_out/r02_nvsmi.log

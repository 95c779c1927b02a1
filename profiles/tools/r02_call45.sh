#!/bin/bash
# round 2, call 45: root backward solve inside the factor kernel, small children issue only their own loads; whole GPU suite
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rxXs > gpurun_out/r02_call45_tests.log 2>&1
python bench.py --no-assoc --no-sharded > gpurun_out/r02_call45_c2.json 2> gpurun_out/r02_call45_c2.err
SLAM_B200_NO_ROOT_FUSE=1 python bench.py --no-assoc --no-sharded > gpurun_out/r02_call45_c2_nofuse.json 2> gpurun_out/r02_call45_c2_nofuse.err
python profiles/tools/front_timeline.py > gpurun_out/r02_call45_timeline.log 2>&1
python profiles/tools/front_timeline.py 1 > gpurun_out/r02_call45_timeline_c1.log 2>&1

#!/bin/bash
# whole GPU suite after the mailbox frame path + warm-up + the patched reference tree; per-frame timing both ways
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rxXs > gpurun_out/r02_call11_tests.log 2>&1
python profiles/tools/frame_path_timing.py > gpurun_out/r02_call11_frame_mailbox.json 2> gpurun_out/r02_call11_frame_mailbox.err
SLAM_B200_FRAME_COPIES=1 python profiles/tools/frame_path_timing.py > gpurun_out/r02_call11_frame_copies.json 2> gpurun_out/r02_call11_frame_copies.err

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rxXs > gpurun_out/r02_call12_tests.log 2>&1

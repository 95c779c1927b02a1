#!/bin/bash
# round 2, call 2 (2 GPUs): guard-band + reproducibility tests, peer-exchange test, default bench line at N=1 and N=2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_peer_exchange_gpu.py -m gpu -q -rxXs -k "guard or reproducible or sanitizer or peer" > gpurun_out/r02_call2_tests.log 2>&1
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02_call2_n1.json 2> gpurun_out/r02_call2_n1.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r02_call2_n2.json 2> gpurun_out/r02_call2_n2.err

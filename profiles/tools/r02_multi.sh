#!/bin/bash
# usage (under gpurun --gpus N): bash profiles/tools/r02_multi.sh N
# the default line exactly as the driver launches it at N GPUs (+ the reference arm) and the 2-GPU tests
N=$1
set -x
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_peer_exchange_gpu.py tests/test_graph_gpu.py -m gpu -q -x -rxXs -k "peer or sharded or guard" > gpurun_out/r02_multi_tests_${N}gpu.log 2>&1
fi
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_default_${N}gpu.json 2> gpurun_out/r02_bench_default_${N}gpu.err ) 2> gpurun_out/r02_bench_default_${N}gpu.time
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_${N}gpu.json 2> gpurun_out/r02_bench_reference_${N}gpu.err

#!/bin/bash
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload c4 --steps 20 --warmup 3 > gpurun_out/r02_c4_${N}gpu.json 2> gpurun_out/r02_c4_${N}gpu.err

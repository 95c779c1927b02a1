# usage: bash profiles/tools/multi_gpu_r01.sh N   (under `gpurun --gpus N`)
N=$1
run() {  # name, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N $2 > gpurun_out/m_$1_${N}gpu.json 2> gpurun_out/m_$1_${N}gpu.err
  echo "$1 rc=$?"; tail -c 400 gpurun_out/m_$1_${N}gpu.json; echo
}
run c2 "--steps 10 --warmup 3 --no-assoc"
run c3 "--workload c3 --steps 5"
run c4 "--workload c4 --steps 20"
run c5 "--workload c5 --steps 10"

# usage: bash profiles/tools/multi_gpu_r01.sh N [workloads]   (under `gpurun --gpus N`; workloads default "c2 c3 c4 c5")
N=$1
W=${2:-"c2 c3 c4 c5"}
run() {  # name, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N $2 > gpurun_out/m_$1_${N}gpu.json 2> gpurun_out/m_$1_${N}gpu.err
  echo "$1 rc=$?"; tail -c 400 gpurun_out/m_$1_${N}gpu.json; echo
}
for w in $W; do
  case $w in
    c2) run c2 "--steps 10 --warmup 3 --no-assoc" ;;
    c3) run c3 "--workload c3 --steps 5" ;;
    c4) run c4 "--workload c4 --steps 20" ;;
    c5) run c5 "--workload c5 --steps 10" ;;
  esac
done

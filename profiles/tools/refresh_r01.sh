set -x
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r_ref.json 2> gpurun_out/r_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r_c2.json 2> gpurun_out/r_c2.err
python bench.py --workload c3 --steps 10 > gpurun_out/r_c3.json 2> gpurun_out/r_c3.err
python bench.py --workload c4 --steps 20 > gpurun_out/r_c4.json 2> gpurun_out/r_c4.err
python bench.py --workload c5 --steps 10 > gpurun_out/r_c5.json 2> gpurun_out/r_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-assoc > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:factor2_kernel --launch-skip 44 --launch-count 11 -f -o gpurun_out/prof_factor_c2 python bench.py --steps 3 --warmup 3 --no-assoc > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assoc_bulk_grid --launch-skip 30 --launch-count 2 -f -o gpurun_out/prof_assoc_c4 python bench.py --workload c4 --steps 4 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none -k regex:assemble_ --launch-skip 20 --launch-count 2 -f -o gpurun_out/prof_asm_c3 python bench.py --workload c3 --steps 3 --replicas 2048 > gpurun_out/ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_ --launch-skip 6 --launch-count 2 -f -o gpurun_out/prof_asm_c5 python bench.py --workload c5 --steps 3 --warmup 3 > gpurun_out/ncu_asm.log 2>&1
ls -la gpurun_out/*.ncu-rep

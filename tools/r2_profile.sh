#!/bin/bash
# Round-2 profile artefacts (one GPU, under gpurun): launch list of the bench command, --set full captures of the dense
# integrate kernel (bench regime and configs[2] real data) and of the hash integrate kernel.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs > $O/r2_prof_bench_plain.json 2> $O/r2_prof_bench_plain.err; echo "plain bench rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_bench_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu --no-configs > $O/r2_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate_dense -s 20 -c 2 -f -o $O/r2_dense_full \
    python bench.py --steps 20 --warmup 5 --no-hash --no-cpu --no-configs > $O/r2_ncu_dense.log 2>&1; echo "ncu dense rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate_hash -s 80 -c 1 -f -o $O/r2_hash_full \
    python bench.py --workload hash --steps 20 --warmup 5 --no-cpu > $O/r2_ncu_hash.log 2>&1; echo "ncu hash rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate_dense -s 60 -c 2 -f -o $O/r2_dense_config2_full \
    python tools/quick_config3.py 0.005 > $O/r2_ncu_config2.log 2>&1; echo "ncu config2 rc=$?"
timeout 300 python tools/quick_config3.py 0.005 2>/dev/null | grep -E "rep|stages" > $O/r2_config2_timing.log; cat $O/r2_config2_timing.log
ls -la $O | grep r2_ | tail

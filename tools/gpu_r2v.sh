#!/bin/bash
# round 2, call v: evidence with the completed-square exp2: bench (both arms), ncu launch list / full capture / DRAM traffic,
# 1M sweep with and without tile skipping, UCV, hill climbing, far rows, per-shape throughput
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_pair \
    python bench.py --steps 1 --warmup 3 --n-test 131072 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pair_kernel -s 6 -c 1 --csv \
    --log-file gpurun_out/r2_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_traffic.log 2>&1
timeout 900 python tools/sweep_bench.py --n 1000000 --dims 1,2,3,4,5,6,7,8 --modes off,on --reps 3 --json gpurun_out/r2_sweep_1m.json > gpurun_out/r2_sweep_1m.log 2>&1; tail -40 gpurun_out/r2_sweep_1m.log
timeout 300 python tools/ucv_bench.py > gpurun_out/r2_ucv_200k.txt 2>&1; tail -3 gpurun_out/r2_ucv_200k.txt
timeout 300 python tools/hc_bench.py --json gpurun_out/r2_hc_config4.json > gpurun_out/hc.log 2>&1; tail -c 400 gpurun_out/hc.log
TUNE_N=400000 python tools/tune_bench.py all 2>&1 | cut -c1-1500

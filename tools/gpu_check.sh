#!/bin/bash
# Runs on the GPU box (under gpurun): GPU tests, the default bench, the ncu launch list and one
# full ncu capture of the pair kernel.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 -f -o gpurun_out/prof_pair \
    python bench.py --steps 1 --warmup 3 --n-test 131072 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pair_kernel -s 3 -c 1 --csv \
    --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_traffic.log 2>&1
ls -la gpurun_out
# per-shape pair-kernel throughput (N = m = 300k), UCV objective, config 4
timeout 200 python tools/tune_bench.py all > gpurun_out/tune_default.log 2>&1
python tools/tune_summary.py gpurun_out/tune_default.log
timeout 100 python tools/ucv_bench.py > gpurun_out/ucv_200k.log 2>&1; tail -3 gpurun_out/ucv_200k.log
timeout 200 python tools/hc_bench.py --json gpurun_out/hc_config4.json > gpurun_out/hc_config4.log 2>&1; tail -c 300 gpurun_out/hc_config4.log

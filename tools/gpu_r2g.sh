#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_skipping_gpu.py tests/test_fullsize_gpu.py tests/test_configs_fullsize_gpu.py tests/test_kde_gpu.py -m gpu -q > gpurun_out/pytest_skip.log 2>&1; echo "skip rc=$?" >> gpurun_out/pytest_skip.log
tail -15 gpurun_out/pytest_skip.log
timeout 600 python tools/sweep_bench.py --n 1000000 --dims 1,2,3,4,6,8 --modes off,on --json gpurun_out/sweep_1m_skip.json > gpurun_out/sweep_1m_skip.log 2>&1; cat gpurun_out/sweep_1m_skip.log
timeout 900 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
# ncu evidence of this round's build: launch list of a bench run, one full capture of the pair kernel, DRAM traffic
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_pair \
    python bench.py --steps 1 --warmup 3 --n-test 131072 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pair_kernel -s 3 -c 1 --csv \
    --log-file gpurun_out/r2_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_traffic.log 2>&1
ls -la gpurun_out | tail -15

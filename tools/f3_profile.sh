#!/bin/bash
# Runs on the GPU box (under gpurun): f3 GPU tests, cdf / sample throughput, the ncu launch list and one full ncu
# capture of the weight kernel (cdf mode and scan mode).  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cdf_sample_gpu.py -q -x > gpurun_out/f3_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/f3_gpu.log
tail -3 gpurun_out/f3_gpu.log
timeout 600 python tools/f3_bench.py 1000000 100000 --cpu > gpurun_out/f3_bench.log 2>&1
tail -10 gpurun_out/f3_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f3_launches.csv \
    python tools/f3_bench.py 1000000 100000 --quick > gpurun_out/f3_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pair_kernel|weight_kernel" -s 1 -c 3 -f -o gpurun_out/prof_f3 \
    python tools/f3_bench.py 1000000 20000 --quick > gpurun_out/f3_ncu_full.log 2>&1
ls -la gpurun_out

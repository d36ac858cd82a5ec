#!/bin/bash
# round 2, second GPU pass: all GPU tests, the far-row / wide-family benchmark, occupancy variants of the f64 pair kernel
set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python tools/far_bench.py > gpurun_out/far_bench.log 2>&1; cat gpurun_out/far_bench.log
timeout 1500 bash tools/tune_all.sh f64 > /dev/null 2>&1; cat gpurun_out/tune.log

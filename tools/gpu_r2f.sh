#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_skipping_gpu.py tests/test_reference_suite_gpu.py tests/test_fullsize_gpu.py tests/test_configs_fullsize_gpu.py -m gpu -q > gpurun_out/pytest_skip.log 2>&1; echo "skip rc=$?" >> gpurun_out/pytest_skip.log
tail -30 gpurun_out/pytest_skip.log
timeout 900 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python tools/sweep_bench.py --n 1000000 --dims 1,2,4,8 --modes off,on --json gpurun_out/sweep_1m_skip.json > gpurun_out/sweep_1m_skip.log 2>&1; cat gpurun_out/sweep_1m_skip.log

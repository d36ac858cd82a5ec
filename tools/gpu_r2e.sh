#!/bin/bash
# round 2, fifth GPU pass: skipping tests first, then everything, the reference suite, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_skipping_gpu.py -m gpu -x -q > gpurun_out/pytest_skip.log 2>&1; echo "skip rc=$?" >> gpurun_out/pytest_skip.log
tail -30 gpurun_out/pytest_skip.log
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err

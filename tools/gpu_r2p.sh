#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_suite_gpu.py -m gpu -q 2>&1 | tail -3
PBN_SKIP_MIN_TRAIN=1000 PBN_SKIP_MIN_TEST=500 PBN_CUDA_WARMUP=0 timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 python tools/memcheck_r2.py > gpurun_out/r2_memcheck.log 2>&1; grep -E "list B|ERROR SUMMARY|MEMCHECK_RUN_OK|Error|error" gpurun_out/r2_memcheck.log | head -20
PBN_SKIP_MIN_TRAIN=1000 PBN_SKIP_MIN_TEST=500 PBN_CUDA_WARMUP=0 timeout 2400 compute-sanitizer --tool racecheck --print-limit 20 python tools/memcheck_r2.py > gpurun_out/r2_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|MEMCHECK_RUN_OK|hazard" gpurun_out/r2_racecheck.log | head -10

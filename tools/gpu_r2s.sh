#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-250
PBN_SKIP_MIN_TRAIN=1000 PBN_SKIP_MIN_TEST=500 PBN_CUDA_WARMUP=0 timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 python tools/memcheck_r2.py > gpurun_out/r2_memcheck.log 2>&1; grep -E "ERROR SUMMARY|MEMCHECK_RUN_OK|Error" gpurun_out/r2_memcheck.log | head
timeout 300 python tools/hc_bench.py --json gpurun_out/r2_hc_config4.json > /dev/null 2>&1; python -c "
import json; d=json.load(open('gpurun_out/r2_hc_config4.json')); print({k:d[k] for k in ('hc_cv_s_per_iter_mean','cache_scores_s','total_s','iterations','final_score')})"

#!/bin/bash
# round 2, call z: sanitizers on the final kernels (completed-square exp2, group skipping, MUFU offload); unroll 8 variant
set -x
mkdir -p gpurun_out
PBN_SKIP_MIN_TRAIN=1000 PBN_SKIP_MIN_TEST=500 PBN_CUDA_WARMUP=0 timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/memcheck_r2.py > gpurun_out/r2_memcheck.log 2>&1; grep -E "ERROR SUMMARY|MEMCHECK_RUN_OK|Error" gpurun_out/r2_memcheck.log | head
PBN_SKIP_MIN_TRAIN=1000 PBN_SKIP_MIN_TEST=500 PBN_CUDA_WARMUP=0 timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/memcheck_r2.py > gpurun_out/r2_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|MEMCHECK_RUN_OK|Error|hazard" gpurun_out/r2_racecheck.log | head
export TUNE_N=400000
python tools/tune_bench.py f64 2>&1 | cut -c1-700
PBN_CUDA_LIB=$PWD/pybnesian_b200/variants/libpbn_u8.so python tools/tune_bench.py f64 2>&1 | cut -c1-700

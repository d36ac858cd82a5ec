#!/bin/bash
# round 2, call t: instruction-class costs next to DFMA; fixed outlier test; int1 / int2 glue variants
set -x
mkdir -p gpurun_out
./tools/micro/fp64_mix > gpurun_out/r2_fp64_mix.txt 2>&1; cat gpurun_out/r2_fp64_mix.txt
timeout 600 python -m pytest tests/test_cv_gpu.py -m gpu -q -x 2>&1 | tail -5
TUNE_N=400000 bash tools/tune_all.sh f64 2>&1 | cut -c1-600
PBN_CUDA_LIB=$PWD/pybnesian_b200/variants/libpbn_int2.so timeout 900 python -m pytest tests/test_kde_gpu.py tests/test_ucv_gpu.py tests/test_cdf_sample_gpu.py -m gpu -q -x 2>&1 | tail -5

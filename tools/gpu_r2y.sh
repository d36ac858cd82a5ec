#!/bin/bash
# round 2, call y: MUFU offload in the packed f32 tile (share of the exponentials on the FP32 pipe), masks 0x80 / 0xC0 / 0x88
set -x
mkdir -p gpurun_out
export TUNE_N=400000 TUNE_SHAPES=ckde:4:float32,kde:1:float32,kde:2:float32,kde:3:float32,kde:4:float32,ckde:2:float32,ckde:3:float32,kde:6:float32
python tools/tune_bench.py 2>&1 | cut -c1-1000
for v in soft0x80 soft0xC0 soft0x88; do PBN_CUDA_LIB=$PWD/pybnesian_b200/variants/libpbn_$v.so python tools/tune_bench.py 2>&1 | cut -c1-1000; done
PBN_CUDA_LIB=$PWD/pybnesian_b200/variants/libpbn_soft0x88.so timeout 900 python -m pytest tests/test_kde_gpu.py tests/test_fullsize_gpu.py tests/test_configs_fullsize_gpu.py tests/test_skipping_gpu.py -m gpu -q -x -k "32 or float32 or f32" 2>&1 | tail -5

#!/bin/bash
# round 2, call w: group skipping inside the units of pass B: correctness (skipping / full-size tests) and throughput
# against the plain pass B, group sizes 2 / per-shape / 4
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_skipping_gpu.py tests/test_configs_fullsize_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -5
export TUNE_N=1000000 TUNE_SKIPPING=1 TUNE_SHAPES=ckde:4:float64,kde:1:float64,kde:2:float64,kde:3:float64,kde:4:float64,ckde:2:float64,ckde:3:float64
echo "== plain pass B"; PBN_GROUP_SKIP=0 python tools/tune_bench.py 2>&1 | cut -c1-900
echo "== group skip, default groups"; python tools/tune_bench.py 2>&1 | cut -c1-900
for v in g2 g4; do echo "== $v"; PBN_CUDA_LIB=$PWD/pybnesian_b200/variants/libpbn_$v.so python tools/tune_bench.py 2>&1 | cut -c1-900; done

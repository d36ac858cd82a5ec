#!/bin/bash
# GPU box: pair-kernel throughput of the default build and every variant under pybnesian_b200/variants/
mkdir -p gpurun_out
: > gpurun_out/tune.log
which=${1:-f64}
python tools/tune_bench.py $which >> gpurun_out/tune.log 2>&1
for f in pybnesian_b200/variants/libpbn_*.so; do
  PBN_CUDA_LIB=$PWD/$f timeout 300 python tools/tune_bench.py $which >> gpurun_out/tune.log 2>&1
done
cat gpurun_out/tune.log

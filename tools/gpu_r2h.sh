#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_skipping_gpu.py tests/test_fullsize_gpu.py tests/test_configs_fullsize_gpu.py tests/test_kde_gpu.py -m gpu -q > gpurun_out/pytest_skip.log 2>&1; echo "skip rc=$?" >> gpurun_out/pytest_skip.log
tail -15 gpurun_out/pytest_skip.log
for lib in default wskip0; do
  if [ $lib = default ]; then unset PBN_CUDA_LIB; else export PBN_CUDA_LIB=$PWD/pybnesian_b200/variants/libpbn_$lib.so; fi
  timeout 600 python tools/sweep_bench.py --n 1000000 --dims 1,2,3,4,6,8 --dtypes float64 --modes off,on > gpurun_out/sweep_1m_$lib.log 2>&1
  echo "== $lib"; cat gpurun_out/sweep_1m_$lib.log
done
unset PBN_CUDA_LIB
timeout 900 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err

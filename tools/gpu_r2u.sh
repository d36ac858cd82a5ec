#!/bin/bash
# round 2, call u: completed-square exp2 (PBN_EXP_SQ): throughput by shape, full GPU suite
set -x
mkdir -p gpurun_out
TUNE_N=400000 python tools/tune_bench.py f64 2>&1 | cut -c1-700
TUNE_N=400000 TUNE_SHAPES=ckde:2:float64,ckde:3:float64,kde:3:float64,kde:5:float64,kde:6:float64,kde:10:float64 python tools/tune_bench.py 2>&1 | cut -c1-800
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-250

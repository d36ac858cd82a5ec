#!/bin/bash
# round 2: does the sorted path + group skipping pay where the boxes prove nothing (families of 6-8 variables, CKDE d = 5..6)?
set -x
export TUNE_N=1000000 TUNE_SKIPPING=1 TUNE_SHAPES=kde:5:float64,kde:6:float64,kde:7:float64,kde:8:float64,ckde:5:float64,ckde:6:float64,ckde:7:float64
echo "== default (fall back above 92% of the units)"; python tools/tune_bench.py 2>&1 | cut -c1-1000
echo "== never fall back"; PBN_SKIP_KEEP_FRAC=1.01 python tools/tune_bench.py 2>&1 | cut -c1-1000

#!/bin/bash
# round 2: do the size thresholds of tile skipping (2^17 training, 2^14 test rows) still make sense with group skipping?
set -x
export TUNE_SHAPES=kde:1:float64,kde:2:float64,kde:3:float64,kde:4:float64,ckde:2:float64,ckde:3:float64,ckde:4:float64,kde:2:float32,ckde:4:float32
for n in 100000 50000; do
  export TUNE_N=$n
  echo "== N = m = $n, all pairs"; TUNE_SKIPPING=0 python tools/tune_bench.py 2>&1 | cut -c1-1100
  echo "== N = m = $n, sorted path forced"; TUNE_SKIPPING=1 PBN_SKIP_MIN_TRAIN=1000 PBN_SKIP_MIN_TEST=1000 python tools/tune_bench.py 2>&1 | cut -c1-1100
done

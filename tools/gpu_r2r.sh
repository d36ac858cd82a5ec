#!/bin/bash
mkdir -p gpurun_out
export TUNE_N=400000
export TUNE_SHAPES="ckde:4:float64,kde:1:float64,kde:2:float64,kde:4:float64,ckde:2:float64,ckde:3:float64"
bash tools/tune_all.sh f64 > /dev/null 2>&1
cat gpurun_out/tune.log
# and once more the default, to see the run-to-run spread
python tools/tune_bench.py f64

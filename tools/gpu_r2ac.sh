#!/bin/bash
# round 2: group skipping with 96 consecutive Morton rows per warp: skipping tests + throughput at 1M x 1M
set -x
timeout 900 python -m pytest tests/test_skipping_gpu.py tests/test_configs_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -3
export TUNE_N=1000000 TUNE_SKIPPING=1 TUNE_SHAPES=ckde:4:float64,kde:2:float64,kde:3:float64,kde:4:float64,ckde:3:float64,ckde:5:float64
python tools/tune_bench.py 2>&1 | cut -c1-900

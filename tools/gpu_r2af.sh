#!/bin/bash
# round 2: group skipping in the packed f32 tile: throughput at 1M x 1M with skipping on, by shape, and the skipping tests
set -x
export TUNE_N=1000000 TUNE_SKIPPING=1 TUNE_SHAPES=kde:1:float32,kde:2:float32,kde:3:float32,kde:4:float32,kde:5:float32,kde:6:float32,ckde:2:float32,ckde:3:float32,ckde:4:float32,ckde:5:float32,ckde:6:float32
echo "== plain pass B"; python tools/tune_bench.py 2>&1 | cut -c1-1400
echo "== group skipping, all shapes"; PBN_GROUP_SKIP_F32=0x1FDFE python tools/tune_bench.py 2>&1 | cut -c1-1400
PBN_GROUP_SKIP_F32=0x1FDFE timeout 900 python -m pytest tests/test_skipping_gpu.py tests/test_configs_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -3

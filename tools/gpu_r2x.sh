#!/bin/bash
# round 2, call x: full GPU suite with group skipping (groups of 2, two or more kernel coordinates); bench; 4 rows per thread for the CKDE shapes
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["value_with_skipping"]["value"], d["value_with_skipping"]["fraction_evaluated"], d["e2e_with_skipping"]["value"], d["roofline"]["frac"], d.get("hc_cv", {}).get("hc_cv_s_per_iter_mean"))
PY
export TUNE_N=400000 TUNE_SHAPES=ckde:4:float64,ckde:2:float64,ckde:3:float64,ckde:5:float64,kde:4:float64,kde:6:float64
python tools/tune_bench.py 2>&1 | cut -c1-900
PBN_CUDA_LIB=$PWD/pybnesian_b200/variants/libpbn_r4.so python tools/tune_bench.py 2>&1 | cut -c1-900

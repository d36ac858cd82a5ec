#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for mode in 0 1; do
  PBN_MORTON_JOINT=$mode timeout 600 python bench.py --no-cpu --no-extras --steps 3 > gpurun_out/bench_morton_$mode.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_morton_$mode.json').read().strip().splitlines()[-1])
w=d['value_with_skipping']; print('PBN_MORTON_JOINT=$mode value',d['value'],'with_skipping',w['value'],w['fraction_evaluated'],w['rel_diff_vs_all_pairs'],'e2e_skip',d['e2e_with_skipping']['value'])
PY
done
timeout 300 python tools/hc_bench.py --json gpurun_out/hc_warm.json > /dev/null 2>&1; python -c "
import json; d=json.load(open('gpurun_out/hc_warm.json')); print({k:d[k] for k in ('hc_cv_s_per_iter_mean','cache_scores_s','setup_s','kernel_warmup_s','total_s','pair_kernel_ms')})"
timeout 300 python -m cProfile -s tottime tools/hc_bench.py 2>&1 | grep -A45 "Ordered by" | cut -c1-150 > gpurun_out/hc_profile.txt; cat gpurun_out/hc_profile.txt

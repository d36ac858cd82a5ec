#!/bin/bash
# round 2, final evidence: GPU suite, bench (both arms), ncu launch list / full capture / DRAM traffic, 1M sweep with and
# without skipping, UCV, hill climbing, far rows, per-shape throughput (default build vs 8 points per step everywhere)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_pair \
    python bench.py --steps 1 --warmup 3 --n-test 131072 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pair_kernel -s 6 -c 1 --csv \
    --log-file gpurun_out/r2_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --e2e-steps 1 > gpurun_out/ncu_traffic.log 2>&1
timeout 900 python tools/sweep_bench.py --n 1000000 --dims 1,2,3,4,5,6,7,8 --modes off,on --reps 3 --json gpurun_out/r2_sweep_1m.json > gpurun_out/r2_sweep_1m.log 2>&1
timeout 300 python tools/ucv_bench.py > gpurun_out/r2_ucv_200k.txt 2>&1; tail -3 gpurun_out/r2_ucv_200k.txt
timeout 300 python tools/hc_bench.py --json gpurun_out/r2_hc_config4.json > gpurun_out/hc.log 2>&1
timeout 600 python tools/far_bench.py > gpurun_out/r2_far_rows_and_wide_families.txt 2>&1; tail -12 gpurun_out/r2_far_rows_and_wide_families.txt
export TUNE_N=400000
python tools/tune_bench.py all 2>&1 | cut -c1-1500
export TUNE_SHAPES=ckde:2:float64,ckde:3:float64,ckde:5:float64,kde:3:float64,kde:5:float64,kde:6:float64
python tools/tune_bench.py 2>&1 | cut -c1-900
PBN_CUDA_LIB=$PWD/pybnesian_b200/variants/libpbn_u8.so python tools/tune_bench.py 2>&1 | cut -c1-900
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["value_with_skipping"]["value"], d["e2e_with_skipping"]["value"], d["roofline"]["frac"], d["roofline"]["issue_model"], d.get("hc_cv", {}).get("hc_cv_s_per_iter_mean"))
print(open("gpurun_out/bench_ref.json").read()[:300])
h = json.load(open("gpurun_out/r2_hc_config4.json")); print({k: h[k] for k in ("hc_cv_s_per_iter_mean", "cache_scores_s", "total_s", "pair_evals_per_s_in_kernel")})
PY

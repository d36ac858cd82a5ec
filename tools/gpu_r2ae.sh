#!/bin/bash
# round 2: sorted path kept for all group-skip-eligible float64 families: GPU suite, bench, 1M sweep
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 900 python tools/sweep_bench.py --n 1000000 --dims 1,2,3,4,5,6,7,8 --modes off,on --reps 3 --json gpurun_out/r2_sweep_1m.json > gpurun_out/r2_sweep_1m.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["value_with_skipping"]["value"], d["e2e_with_skipping"]["value"], d["roofline"]["frac"], d.get("hc_cv", {}).get("hc_cv_s_per_iter_mean"))
for r in json.load(open("gpurun_out/r2_sweep_1m.json")):
    print(r["dtype"][-2:], r["d"], r["tile_skipping"], "%.3e" % r["pair_evals_per_s_job_wall"], round(r.get("units_evaluated_fraction", 1), 3))
PY

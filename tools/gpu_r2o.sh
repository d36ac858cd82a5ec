#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/reference_suite.json"))
print("REFERENCE SUITE: collected", d["collected"], "passed", d["passed"], "skipped", d["skipped"], "failed", len(d["failed"]))
for k, v in d["failed"].items():
    print("  FAIL", k, "|", v.replace("\n", " ")[:200])
PY
PBN_SKIP_MIN_TRAIN=1000 PBN_SKIP_MIN_TEST=500 timeout 300 python tools/memcheck_r2.py > gpurun_out/r2_memcheck_plain.log 2>&1; tail -4 gpurun_out/r2_memcheck_plain.log
PBN_SKIP_MIN_TRAIN=1000 PBN_SKIP_MIN_TEST=500 PBN_CUDA_WARMUP=0 timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/memcheck_r2.py > gpurun_out/r2_memcheck.log 2>&1; tail -12 gpurun_out/r2_memcheck.log

#!/bin/bash
# round 2, third GPU pass: all GPU tests (no -x), incl. the reference's own test files (staged copy)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/reference_suite.json"))
    print("REFERENCE SUITE: collected", d["collected"], "passed", d["passed"], "skipped", d["skipped"], "failed", len(d["failed"]))
    for k, v in d["failed"].items():
        print("  FAIL", k, "|", v.replace("\n", " ")[:200])
except Exception as e:
    print("no reference suite summary:", e)
PY

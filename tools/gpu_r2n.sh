#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_reference_suite_gpu.py -m gpu -q > gpurun_out/pytest_ref.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_ref.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/reference_suite.json"))
print("REFERENCE SUITE: collected", d["collected"], "passed", d["passed"], "skipped", d["skipped"], "failed", len(d["failed"]))
for k, v in d["failed"].items():
    print("  FAIL", k, "|", v.replace("\n", " ")[:260])
PY

#!/bin/bash
# round 2, 2-GPU box, final build: the two tests that need 2 devices, bench under torchrun at N = 2
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests/test_cv_gpu.py -m gpu -q -k "multi_gpu or multi_device" > gpurun_out/pytest_2gpu.log 2>&1; tail -4 gpurun_out/pytest_2gpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_2gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["value_with_skipping"]["value"])
print("strong", d.get("strong_scaling")); print("hc", {k: v for k, v in d.get("hc_cv", {}).items() if k != "operators"})
print("inproc", d.get("inproc_multi_gpu"))
PY

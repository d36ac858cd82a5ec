#!/bin/bash
# round 2 (2 GPUs): both multi-GPU tests, the in-process check with timing, bench at N = 2 under torchrun
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_cv_gpu.py -m gpu -q -k "multi" > gpurun_out/pytest_multi_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi_2gpu.log
tail -8 gpurun_out/pytest_multi_2gpu.log
timeout 900 python tools/inproc_multi_gpu_check.py --json gpurun_out/r2_inproc_multi_gpu_2.json > gpurun_out/r2_inproc_multi_gpu_2.log 2>&1; tail -3 gpurun_out/r2_inproc_multi_gpu_2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
cat gpurun_out/r2_bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_2gpu.json 2>/dev/null; cat gpurun_out/r2_bench_reference_2gpu.json | cut -c1-400

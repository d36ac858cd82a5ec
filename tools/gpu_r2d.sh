#!/bin/bash
# round 2, fourth GPU pass (2 GPUs): all GPU tests incl. both multi-GPU tests and the reference suite, the in-process
# multi-device check with timing, bench.py at N = 2 under torchrun
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_2gpu.log
tail -40 gpurun_out/pytest_gpu_2gpu.log
timeout 900 python tools/inproc_multi_gpu_check.py --json gpurun_out/inproc_multi_gpu_2.json > gpurun_out/inproc_multi_gpu_2.log 2>&1; tail -5 gpurun_out/inproc_multi_gpu_2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
cat gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err

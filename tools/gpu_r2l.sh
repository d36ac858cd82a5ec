#!/bin/bash
# round 2 (8 GPUs): in-process multi-device check, bench.py at N = 8 under torchrun (incl. its strong-scaling, HC and
# in-process legs), config-5 sweep at 16M rows from ONE process over all 8 devices
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python tools/inproc_multi_gpu_check.py --no-timing --json gpurun_out/r2_inproc_multi_gpu_8.json > gpurun_out/r2_inproc_multi_gpu_8.log 2>&1; tail -2 gpurun_out/r2_inproc_multi_gpu_8.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench rc=$?"
cat gpurun_out/r2_bench_8gpu.json | cut -c1-6000; tail -3 gpurun_out/bench_8gpu.err
timeout 1500 python tools/sweep_bench.py --n 16000000 --dims 1,4 --modes off,on --json gpurun_out/r2_sweep_16m_8gpu_d14.json > gpurun_out/r2_sweep_16m_8gpu_d14.log 2>&1; cat gpurun_out/r2_sweep_16m_8gpu_d14.log
timeout 900 python tools/sweep_bench.py --n 16000000 --dims 8 --modes off --json gpurun_out/r2_sweep_16m_8gpu_d8.json > gpurun_out/r2_sweep_16m_8gpu_d8.log 2>&1; cat gpurun_out/r2_sweep_16m_8gpu_d8.log

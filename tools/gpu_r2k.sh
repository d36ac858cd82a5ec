#!/bin/bash
# round 2: config-5 sweep at 8M and 16M rows on ONE GPU (all pairs, and with tile skipping where it can skip)
set -x
mkdir -p gpurun_out
timeout 1500 python tools/sweep_bench.py --n 8000000 --dims 1,4,8 --modes off,on --json gpurun_out/r2_sweep_8m.json > gpurun_out/r2_sweep_8m.log 2>&1; cat gpurun_out/r2_sweep_8m.log
timeout 2000 python tools/sweep_bench.py --n 16000000 --dims 1,4 --modes off,on --json gpurun_out/r2_sweep_16m_d14.json > gpurun_out/r2_sweep_16m_d14.log 2>&1; cat gpurun_out/r2_sweep_16m_d14.log
timeout 1200 python tools/sweep_bench.py --n 16000000 --dims 8 --modes off --json gpurun_out/r2_sweep_16m_d8.json > gpurun_out/r2_sweep_16m_d8.log 2>&1; cat gpurun_out/r2_sweep_16m_d8.log

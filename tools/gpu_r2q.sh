#!/bin/bash
# the whole GPU suite on a 2-GPU box: the default context is then the multi-device one
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_2gpu_default_ctx.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_2gpu_default_ctx.log
tail -25 gpurun_out/pytest_gpu_2gpu_default_ctx.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

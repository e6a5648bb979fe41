#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests + smoke; logs land in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -k "${PYTEST_K:-test}" 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log

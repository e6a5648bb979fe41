#!/bin/bash
# 2-GPU: NCCL test with the symmetric-memory exchange, timeline; 1-GPU: renderer / scene / stove tests, multiball profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py tests/test_gpu_scene.py tests/test_gpu_stove.py -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_i.log
echo "pytest exit: ${PIPESTATUS[0]}"; tail -8 gpurun_out/pytest_i.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 scripts/timeline_dp.py > gpurun_out/timeline_dp.log 2>&1
echo "timeline exit $?"; tail -2 gpurun_out/timeline_dp.log | cut -c1-300
timeout 300 python scripts/profile_variant.py o6 > gpurun_out/profile_o6.log 2>&1; echo "profile exit $?"; head -30 gpurun_out/profile_o6.txt

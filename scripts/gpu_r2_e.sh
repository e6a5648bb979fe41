#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 scripts/timeline_dp.py > gpurun_out/timeline_dp.log 2>&1
echo "timeline exit $?"; tail -3 gpurun_out/timeline_dp.log
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_stove.py -m gpu -q --timeout 300 2>&1 | tail -8

#!/bin/bash
# 2-GPU call: NCCL parity test (overlapped exchange), parametric dynamics test, bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp_nccl.py tests/test_gpu_dynamics.py tests/test_gpu_encoder.py -m gpu -q --timeout 300 2>&1 | tail -25 > gpurun_out/pytest_dp.log
echo "pytest exit: ${PIPESTATUS[0]}"; tail -25 gpurun_out/pytest_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench2 exit $?"; head -c 1200 gpurun_out/bench_2gpu.json; echo; tail -5 gpurun_out/bench_2gpu.err

#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/microbench.py > gpurun_out/microbench.log 2>&1; echo "microbench exit $?"; tail -2 gpurun_out/microbench.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 scripts/allreduce_bench.py > gpurun_out/allreduce_bench.log 2>&1
echo "allreduce exit $?"; grep floats gpurun_out/allreduce_bench.log | cut -c1-1500; tail -3 gpurun_out/allreduce_bench.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_stove.py -m gpu -q --timeout 600 -k "other_baseline or golden" 2>&1 | tail -6

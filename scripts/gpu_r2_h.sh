#!/bin/bash
# 8-GPU call: all-reduce latency at N = 8, the bench at N = 8 (scaling), DP timeline of rank 0
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 scripts/allreduce_bench.py > gpurun_out/allreduce_bench8.log 2>&1
echo "allreduce exit $?"; grep floats gpurun_out/allreduce_bench8.log | cut -c1-900
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 scripts/timeline_dp.py > gpurun_out/timeline_dp8.log 2>&1
echo "timeline exit $?"; cp gpurun_out/timeline_dp.txt gpurun_out/timeline_dp8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
echo "bench8 exit $?"; head -c 400 gpurun_out/bench_8gpu.json; echo; tail -3 gpurun_out/bench_8gpu.err

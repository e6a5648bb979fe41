#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 scripts/timeline_dp.py > gpurun_out/timeline_dp8.log 2>&1
echo "timeline exit $?"; cp gpurun_out/timeline_dp.txt gpurun_out/timeline_dp8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
echo "bench8 exit $?"; head -c 1600 gpurun_out/bench_8gpu.json; echo; tail -3 gpurun_out/bench_8gpu.err

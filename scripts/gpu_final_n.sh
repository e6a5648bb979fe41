#!/bin/bash
# end-of-round scaling check on N GPUs (N = number of visible devices): reference arm is rank-0-only, then the bench
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench N=$N exit $?"; head -c 400 gpurun_out/bench_${N}gpu.json; echo; tail -2 gpurun_out/bench_${N}gpu.err
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q 2>&1 | tail -2; fi

#!/bin/bash
# round-2 call A: new LSTM kernels (checks + timings), then the GPU suite and a bench line
mkdir -p gpurun_out
timeout 300 python scripts/lstm_tc_bench.py > gpurun_out/lstm_tc_bench.log 2>&1
echo "lstm bench exit $?"; tail -40 gpurun_out/lstm_tc_bench.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}"; tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}"; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref exit $?"; tail -c 1500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err

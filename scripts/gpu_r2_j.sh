#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}"; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; head -c 600 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
timeout 300 python scripts/timeline_graph.py > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
bash scripts/gpu_sanitizer.sh

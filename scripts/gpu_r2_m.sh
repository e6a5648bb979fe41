#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; head -c 300 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
timeout 300 python scripts/timeline_graph.py > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL='tests/test_gpu_scene.py tests/test_gpu_spn.py tests/test_gpu_stove.py -k golden -m gpu -q -x -p no:cacheprovider'
echo "--- racecheck sequence (as before)"
timeout 600 $SAN --tool racecheck --print-limit 5 python -m pytest $SEL 2>&1 | grep -E "passed|failed|RACECHECK|FAILED" | cut -c1-300
echo "--- racecheck sequence, CUDA_LAUNCH_BLOCKING=1"
CUDA_LAUNCH_BLOCKING=1 timeout 600 $SAN --tool racecheck --print-limit 5 python -m pytest $SEL 2>&1 | grep -E "passed|failed|RACECHECK|FAILED" | cut -c1-300
echo "--- racecheck sequence, no caching allocator"
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 600 $SAN --tool racecheck --print-limit 5 python -m pytest $SEL 2>&1 | grep -E "passed|failed|RACECHECK|FAILED" | cut -c1-300

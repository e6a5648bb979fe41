#!/bin/bash
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
echo "--- racecheck, envs alone"
timeout 600 $SAN --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_stove.py -k "test_stove_golden and envs" -m gpu -q -x -p no:cacheprovider 2>&1 | grep -E "passed|failed|RACECHECK|parity failures" | cut -c1-400
echo "--- racecheck, plain alone"
timeout 600 $SAN --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_stove.py -k "test_stove_golden and plain and not generic and not nw1" -m gpu -q -x -p no:cacheprovider 2>&1 | grep -E "passed|failed|RACECHECK|parity failures" | cut -c1-400
echo "--- initcheck, envs alone"
timeout 600 $SAN --tool initcheck --print-limit 10 python -m pytest tests/test_gpu_stove.py -k "test_stove_golden and envs" -m gpu -q -x -p no:cacheprovider > gpurun_out/initcheck_envs.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|Uninitialized|parity failures" gpurun_out/initcheck_envs.log | head -20 | cut -c1-300
grep -E "Uninitialized" -A12 gpurun_out/initcheck_envs.log | grep -E "Uninitialized|in .*kernel|at .*cu" | head -40 | cut -c1-200
echo "--- memcheck, envs alone"
timeout 600 $SAN --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_stove.py -k "test_stove_golden and envs" -m gpu -q -x -p no:cacheprovider 2>&1 | grep -E "passed|failed|ERROR SUMMARY|parity failures" | cut -c1-400
echo "--- CUDA_LAUNCH_BLOCKING, envs alone"
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_stove.py -k "test_stove_golden and envs" -m gpu -q -x -p no:cacheprovider 2>&1 | grep -E "passed|failed|parity failures" | cut -c1-400

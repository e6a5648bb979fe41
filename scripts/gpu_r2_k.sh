#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; head -c 300 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
timeout 300 python scripts/timeline_graph.py > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $SAN --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_scene.py tests/test_gpu_spn.py tests/test_gpu_stove.py -k "golden" -m gpu -q -x --timeout 550 -p no:cacheprovider > gpurun_out/sanitizer_racecheck_scene_spn_stove.log 2>&1
echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_racecheck_scene_spn_stove.log | tail -3

#!/bin/bash
# compute-sanitizer over the GPU parity tests: memcheck on every kernel family, racecheck on the
# kernels with named barriers / mbarriers / cp.async pipelines.  Summaries land in gpurun_out/sanitizer_*.txt
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {   # name tool tests...
  local name=$1 tool=$2; shift 2
  timeout ${SAN_TIMEOUT:-900} $SAN --tool $tool --print-limit 20 --error-exitcode 77 \
      python -m pytest "$@" -m gpu -q -x --timeout 850 -p no:cacheprovider > gpurun_out/sanitizer_${name}.log 2>&1
  local rc=$?
  { echo "== $name: compute-sanitizer --tool $tool pytest $* -> exit $rc"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitizer_${name}.log | tail -6; } >> gpurun_out/sanitizer_summary.txt
}
rm -f gpurun_out/sanitizer_summary.txt
run memcheck_encoder memcheck tests/test_gpu_encoder.py
run memcheck_scene_spn memcheck tests/test_gpu_scene.py tests/test_gpu_spn.py
run memcheck_dynamics_optim memcheck tests/test_gpu_dynamics.py tests/test_gpu_optim.py
run memcheck_stove memcheck tests/test_gpu_stove.py -k "golden or quickstart or in_place"
run racecheck_encoder racecheck tests/test_gpu_encoder.py -k "tc3_gemm or 129-36 or 5-1024 or head"
run racecheck_dynamics racecheck tests/test_gpu_dynamics.py -k "golden"
run racecheck_scene_spn racecheck tests/test_gpu_scene.py tests/test_gpu_spn.py -k "golden"
cat gpurun_out/sanitizer_summary.txt

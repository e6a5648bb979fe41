#!/bin/bash
# compute-sanitizer over the fused scene-likelihood kernels (scene_ll.cu / scene_ll_bwd.cu): memcheck on every case of
# tests/test_gpu_scene_ll.py, racecheck (shared-memory hazards: the kernels alias regions across phases, exchange tiles
# between warps and use shared atomics) on the small cases.  Summary -> gpurun_out/sanitizer_scene_ll_summary.txt
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
OUT=gpurun_out/sanitizer_scene_ll_summary.txt
rm -f $OUT
run() {
  local name=$1 tool=$2; shift 2
  timeout ${SAN_TIMEOUT:-900} $SAN --tool $tool --print-limit 20 --error-exitcode 77 \
      python -m pytest "$@" -m gpu -q -x --timeout 850 -p no:cacheprovider > gpurun_out/sanitizer_${name}.log 2>&1
  local rc=$?
  { echo "== $name: compute-sanitizer --tool $tool pytest $* -> exit $rc"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitizer_${name}.log | tail -6; } >> $OUT
}
run memcheck_scene_ll memcheck tests/test_gpu_scene_ll.py
run racecheck_scene_ll racecheck tests/test_gpu_scene_ll.py -k "kw0-1 or kw1-5 or kw2-33 or kw5-40 or kw6-70 or kw8-7 or oracle"
run racecheck_scene_seq racecheck tests/test_gpu_scene_ll.py -k "sequence and (kw0-5 or kw2-9)"
run initcheck_scene_ll initcheck tests/test_gpu_scene_ll.py -k "kw2-33 or kw6-70 or (sequence and kw0-5)"
cat $OUT

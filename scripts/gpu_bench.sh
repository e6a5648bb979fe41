#!/bin/bash
# On the GPU box: bench line + ncu launch list of one training step (+ a 200-step rollout).
mkdir -p gpurun_out
timeout 900 python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 1 --warmup 3 --rollout 200 > gpurun_out/ncu_list.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv

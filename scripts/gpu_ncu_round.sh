#!/bin/bash
# One GPU call: launch list of an eager training step (+ 200-step rollout) and `ncu --set full` of the named kernels.
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 1 --warmup 3 --rollout 200 > gpurun_out/ncu_list.log 2>&1
echo "list exit $?"; wc -l gpurun_out/launches.csv
for K in ${KERNELS}; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:$K -c ${COUNT:-1} -o gpurun_out/full_$K -f \
      python scripts/profile_step.py --steps 1 --warmup 3 > gpurun_out/ncu_$K.log 2>&1
  echo "$K exit $?"
done
ls -la gpurun_out/*.ncu-rep

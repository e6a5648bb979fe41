#!/bin/bash
# ncu --set full, one launch per named kernel (reports must stay small: gpurun_out <= 64 MiB)
mkdir -p gpurun_out
for K in ${KERNELS:-gnn_fwd_kernel gnn_bwd_kernel}; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:$K -c 1 -o gpurun_out/full_$K -f \
      python scripts/profile_step.py --steps 1 --warmup 3 --rollout ${ROLL:-0} > gpurun_out/ncu_$K.log 2>&1
  echo "$K exit $?"
done
ls -la gpurun_out/*.ncu-rep

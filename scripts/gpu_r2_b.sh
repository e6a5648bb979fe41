#!/bin/bash
# round-2 call B: GPU suite, graph-replay timeline, ncu --set full of the LSTM kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python scripts/timeline_graph.py > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
for K in ${KERNELS:-tc3_gemm lstm_cell_bwd_t lstm_gemm_cell_fwd}; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:$K -c ${COUNT:-4} -o gpurun_out/full_$K -f \
      python scripts/profile_step.py --steps 1 --warmup 3 > gpurun_out/ncu_$K.log 2>&1
  echo "$K exit $?"
done
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 1" "1 2" "2 2" "2 3" "2 4" "4 4" "3 3"; do
  set -- $cfg
  STOVE_GNN_SEQ_FWD=$1 STOVE_GNN_SEQ_BWD=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']['kernel_ms_per_step']
print('fwd=$1 bwd=$2', round(d['ms_per_step'],3), 'dynstep_fwd', r.get('dynstep_fwd'), 'dynstep_bwd', r.get('dynstep_bwd'))" | tee -a gpurun_out/sweep.txt
done

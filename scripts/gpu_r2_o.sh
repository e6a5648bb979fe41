#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_stove.py -m gpu -q --timeout 600 2>&1 | tail -4
timeout 300 python scripts/tune_head_par.py 2>&1 | grep head_par | tee gpurun_out/tune_head_par.txt
timeout 120 python scripts/lstm_tc_bench.py --skip-check 2>&1 | tail -20 | tee gpurun_out/lstm_tc_bench.log

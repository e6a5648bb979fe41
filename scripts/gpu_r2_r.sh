#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/tune_head_par.py wgrad_ctas 2>&1 | grep wgrad_ctas | tee gpurun_out/tune_wgrad_ctas.txt

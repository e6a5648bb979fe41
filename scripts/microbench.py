"""Measured roofline denominators (run on the GPU box): FP32 FMA peak and tcgen05 TF32 issue peak.
Writes gpurun_out/microbench.json (copy to profiles/r02_microbench.json)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stove_b200 import _native as N

lib = N.lib()
scratch = torch.zeros(16, device='cuda')
out = {'gpu': torch.cuda.get_device_name(0)}
v = C.c_double()
best = 0.0
for ctas in (148, 296, 592):
    for _ in range(3):
        N.check(lib.stove_microbench_ffma(ctas, 20000, N.ptr(scratch), C.byref(v), N.stream()))
        best = max(best, v.value)
out['fp32_fma_tflops'] = best
out['fp32_fma_how'] = 'best of 9 runs: 148/296/592 CTAs x 1024 threads x 20000 x 16 independent FFMA, CUDA events'
for n_cols in (128, 256):
    best = 0.0
    for _ in range(3):
        N.check(lib.stove_microbench_tf32(148, 20000, n_cols, C.byref(v), N.stream()))
        best = max(best, v.value)
    out['tcgen05_tf32_tflops_n%d' % n_cols] = best
out['tcgen05_tf32_tflops'] = max(out['tcgen05_tf32_tflops_n128'], out['tcgen05_tf32_tflops_n256'])
out['tcgen05_tf32_how'] = ('best of 3: 148 CTAs, one elected thread each issuing 80000 tcgen05.mma.cta_group::1.kind::tf32 '
                           '128 x N x 8 on resident shared-memory tiles (SS mode), CUDA events; sustained ~0.5 s runs')
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/microbench.json', 'w') as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))

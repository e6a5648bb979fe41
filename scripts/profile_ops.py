"""ATen-level view of one eager training step (torch.profiler): which library ops still launch kernels
around the native ones.  Prints ops sorted by CUDA time with their launch counts."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from stove_b200 import dp  # noqa: E402

dev = torch.device('cuda', 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model = bench.build_model(dev)
eng = dp.DataParallel(model)
xs = [bench.make_frames(bench.BATCH, i).to(dev) for i in range(2)]
for i in range(3):
    eng.forward_backward(xs[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    eng.forward_backward(xs[0])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=70, max_name_column_width=60))

"""Profiling target for ncu (--profile-from-start off): one eager training step of a bench variant between
cudaProfilerStart/Stop.   ncu ... python scripts/profile_variant_ncu.py o6 [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stove_b200 import dp, synth

variant = sys.argv[1] if len(sys.argv) > 1 else 'o6'
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device('cuda', 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model = bench.build_variant(variant, dev, seed=4)
c = model.c
x = bench.make_frames(batch, 3, num_obj=min(c.num_obj, 6) if c.num_obj > 3 else 3, res=c.width).to(dev)
a = synth.random_actions(batch, 8, 9, 1).to(dev) if c.action_conditioned else None
eng = dp.DataParallel(model)
for _ in range(3):
    eng.forward_backward(x, 1, a)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.forward_backward(x, 1, a)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()

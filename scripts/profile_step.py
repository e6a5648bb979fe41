"""Profiling target: W warm-up training steps, then K steps between cudaProfilerStart/Stop
(use with `ncu --profile-from-start off`).  Optionally also one long rollout."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=1)
ap.add_argument('--warmup', type=int, default=3)
ap.add_argument('--rollout', type=int, default=0, help='also profile one AC rollout of this many steps')
a = ap.parse_args()
from stove_b200 import dp, synth  # noqa: E402
dev = torch.device('cuda', 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model = bench.build_model(dev)
eng = dp.DataParallel(model)
xs = [bench.make_frames(bench.BATCH, i).to(dev) for i in range(2)]
for i in range(a.warmup):
    eng.forward_backward(xs[i % 2])
if a.rollout:
    ac = bench.build_ac_model(dev)
    g = torch.Generator().manual_seed(0)
    zl = torch.cat([0.2 + 0.3 * torch.rand(1024, 3, 2, generator=g), torch.rand(1024, 3, 16, generator=g) - 0.5], -1).to(dev)
    app = torch.rand(1024, 3, 3, generator=g).to(dev)
    act = synth.random_actions(1024, a.rollout, 9, 1).to(dev)
    ac.rollout(zl, 4, actions=act, appearance=app)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for i in range(a.steps):
    eng.forward_backward(xs[i % 2])
if a.rollout:
    ac.rollout(zl, a.rollout, actions=act, appearance=app)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')

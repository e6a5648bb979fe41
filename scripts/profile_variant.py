"""Per-kernel device time of one eager training step of a bench variant (library event profiler, side streams
off): python scripts/profile_variant.py o6 [batch] [option=value ...]   -> gpurun_out/profile_<variant>.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stove_b200 import _native as N, dp, ops, synth

args = [a for a in sys.argv[1:] if '=' not in a]
for a in sys.argv[1:]:
    if '=' in a:                                   # library options, e.g. gnn_threads=512
        k, v = a.split('=')
        N.set_option(k, int(v))
variant = args[0] if len(args) > 0 else 'o6'
batch = int(args[1]) if len(args) > 1 else 256
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
ops.set_fork(False)
lib = N.lib()
lib.stove_profile_enable(1)
N.profile_read()
K = 5
for _ in range(K):
    eng.forward_backward(x, 1, a)
torch.cuda.synchronize()
lib.stove_profile_enable(0)
ops.set_fork(True)
per = {}
for name, t in N.profile_read():
    per.setdefault(name, []).append(t)
rows = sorted(((sum(v) / K, len(v) / K, k) for k, v in per.items()), reverse=True)
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/profile_%s.txt' % variant, 'w') as f:
    f.write('# %s, batch %d: native kernels, ms per step (serial, side streams off), launches per step\n' % (variant, batch))
    f.write('# total %.3f ms\n' % sum(r[0] for r in rows))
    for ms, cnt, k in rows:
        f.write('%8.4f %6.1f  %s\n' % (ms, cnt, k))
print(open('gpurun_out/profile_%s.txt' % variant).read())

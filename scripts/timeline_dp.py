"""Timeline of one replay of the captured data-parallel training step on rank 0 (launch with torch.distributed.run,
>= 2 ranks): start offset, duration and stream of every kernel incl. the NCCL ones, to read the overlap of the
piecewise gradient exchange with the LSTM backward.  Output: gpurun_out/timeline_dp.txt"""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from stove_b200 import dp  # noqa: E402

world, rank, local = bench.dist_setup(int(os.environ.get('WORLD_SIZE', '1')))
dev = torch.device('cuda', local)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model = bench.build_model(dev)
eng = dp.DataParallel(model)
xs = [bench.make_frames(bench.BATCH, 10 * rank + i).to(dev) for i in range(2)]
g = dp.GraphedStep(eng, xs[0])
for i in range(5):
    g(xs[i % 2])
torch.cuda.synchronize()
import torch.distributed as dist
if world > 1:
    dist.barrier()
if rank == 0:
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(3):
            g(xs[i % 2])
        torch.cuda.synchronize()
    os.makedirs('gpurun_out', exist_ok=True)
    prof.export_chrome_trace('gpurun_out/trace_dp.json')
    ev = [e for e in json.load(open('gpurun_out/trace_dp.json'))['traceEvents']
          if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'ts' in e]
    ev.sort(key=lambda e: e['ts'])
    first = [i for i, e in enumerate(ev) if 'bw_transform' in e['name']]
    lo = first[-1]                      # the last replay
    last = ev[lo:]
    t0 = last[0]['ts']
    with open('gpurun_out/timeline_dp.txt', 'w') as f:
        f.write('# last replay of the captured DP training step on rank 0 of %d: %d GPU activities\n' % (world, len(last)))
        f.write('# start_us  dur_us  stream  name\n')
        for e in last:
            f.write('%8.1f %7.1f  %4s  %s\n' % (e['ts'] - t0, e['dur'], e.get('args', {}).get('stream', '?'), e['name'][:90]))
    os.remove('gpurun_out/trace_dp.json')
else:
    for i in range(3):
        g(xs[i % 2])
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
os._exit(0)

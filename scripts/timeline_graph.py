"""Timeline of one REPLAY of the captured training step (CUPTI kernel records through torch.profiler):
start offset, duration and stream of every kernel, so the critical chain and the idle gaps of the graph can
be read off.  Output: gpurun_out/timeline.txt"""
import json
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
g = dp.GraphedStep(eng, xs[0])
for i in range(5):
    g(xs[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        g(xs[i % 2])
    torch.cuda.synchronize()
os.makedirs('gpurun_out', exist_ok=True)
prof.export_chrome_trace('gpurun_out/trace.json')
ev = [e for e in json.load(open('gpurun_out/trace.json'))['traceEvents']
      if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'ts' in e]
ev.sort(key=lambda e: e['ts'])
# split replays at large gaps
groups, cur = [], [ev[0]]
for a, b in zip(ev, ev[1:]):
    if b['ts'] - (a['ts'] + a['dur']) > 30 and b['name'].startswith('Memcpy'):
        groups.append(cur)
        cur = []
    cur.append(b)
groups.append(cur)
last = groups[-1]
t0 = last[0]['ts']
end = max(e['ts'] + e['dur'] for e in last)
with open('gpurun_out/timeline.txt', 'w') as f:
    f.write('# one replay of the captured training step: %d GPU activities, %.1f us wall\n' % (len(last), end - t0))
    f.write('# start_us  dur_us  stream  name\n')
    for e in last:
        f.write('%8.1f %7.1f  %4s  %s\n' % (e['ts'] - t0, e['dur'], e.get('args', {}).get('stream', '?'), e['name'][:90]))
print(open('gpurun_out/timeline.txt').read()[:200])
os.remove('gpurun_out/trace.json')

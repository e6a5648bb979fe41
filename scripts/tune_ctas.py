"""Step time of the captured config-1 training step for several CTA caps of a side-stream kernel
(it runs beside the LSTM backward and cannot share an SM with the GEMM CTAs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stove_b200 import _native as N, dp

dev = torch.device('cuda', 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
pool = [bench.make_frames(bench.BATCH, i).to(dev) for i in range(bench.POOL)]
OPT = sys.argv[1] if len(sys.argv) > 1 else 'head_par_ctas'
for cap in (0, 111, 96, 74, 64, 48):
    N.set_option(OPT, cap)
    model = bench.build_model(dev)
    eng = dp.DataParallel(model)
    g = dp.GraphedStep(eng, pool[0])
    ms, reps = bench.timed(lambda i: g(pool[i % bench.POOL]), 20, 5, 1)
    print('%s %3d: %.4f ms/step (median of %d regions)' % (OPT, cap, ms / 20, reps), flush=True)
    del g

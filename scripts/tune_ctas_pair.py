"""Step time of the captured config-1 training step for pairs of CTA caps (dynloop_wgrad, head_bwd_par): with the
dynamics weights packed on a stream of their own both kernels run beside the first GEMM of the LSTM backward, whose
CTAs need a whole SM's shared memory.  One process, one box: the lines are comparable with each other."""
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
PAIRS = [(74, 0), (48, 0), (74, 64), (48, 64), (37, 0), (60, 0), (48, 48), (74, 48), (37, 64), (96, 0), (74, 96), (30, 0)]
for wg, hp in PAIRS:
    N.set_option('wgrad_ctas', wg)
    N.set_option('head_par_ctas', hp)
    model = bench.build_model(dev)
    eng = dp.DataParallel(model)
    g = dp.GraphedStep(eng, pool[0])
    ms, reps = bench.timed(lambda i: g(pool[i % bench.POOL]), 20, 5, 1)
    print('wgrad_ctas %3d head_par_ctas %3d: %.4f ms/step (median of %d regions)' % (wg, hp, ms / 20, reps), flush=True)
    del g

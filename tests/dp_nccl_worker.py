"""Worker of tests/test_gpu_dp_nccl.py (one process per GPU, launched through torch.distributed.run): the
all-reduced gradient of rank-local shards of the real Stove model equals the single-rank gradient of the
concatenated batch, for the plain and for the overlapped (piecewise) exchange.  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    world, rank, local = bench.dist_setup(int(os.environ.get('WORLD_SIZE', '1')))
    dev = torch.device('cuda', local)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = bench.build_model(dev)
    import torch.distributed as dist
    for p in model.parameters():
        dist.broadcast(p.data, 0)
    out = bench.dp_gradient_check(model, dev, world, rank)
    # the same through a captured graph: replay == eager loss on the same batch
    from stove_b200 import dp
    eng = dp.DataParallel(model, broadcast=False)
    x = bench.make_frames(16, 99 + rank).to(dev)
    g = dp.GraphedStep(eng, x)
    l1 = float(g(x))
    l2 = float(g(x))
    out['graph_replay_finite'] = l1 == l1 and l2 == l2
    out['graph_overlap_active'] = eng._overlap is not None
    torch.cuda.synchronize()
    if rank == 0:
        print('DPCHECK ' + json.dumps(out), flush=True)
    dist.barrier()
    os._exit(0)


if __name__ == '__main__':
    main()

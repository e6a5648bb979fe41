"""All-reduce latency at the gradient-bucket sizes (run with torch.distributed.run, >= 2 ranks): NCCL vs the
symmetric-memory kernels of torch (one-shot / two-shot / multimem), eager and replayed from a CUDA graph.
Output: gpurun_out/allreduce_bench_N<world>.json on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

world, rank, local = bench.dist_setup(int(os.environ.get('WORLD_SIZE', '1')))
dev = torch.device('cuda', local)
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

group = dist.group.WORLD
out = {'world': world}


def time_it(fn, iters=50):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def graphed(fn, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(iters):
            fn()
    return time_it(g.replay, iters=10) / iters


for numel in (350_000, 525_000, 1_410_255 // 4 * 4, 2_000_000):
    key = '%d_floats' % numel
    res = {}
    x = torch.randn(numel, device=dev)
    res['nccl_eager_us'] = time_it(lambda: dist.all_reduce(x))
    try:
        res['nccl_graph_us'] = graphed(lambda: dist.all_reduce(x))
    except Exception as e:      # noqa: BLE001
        res['nccl_graph_us'] = repr(e)[:200]
    try:
        t = symm_mem.empty(numel, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(t, group=group.group_name)
        t.normal_()
        res['multicast_ptr'] = int(hdl.multicast_ptr != 0)
        ref = t.clone()
        dist.all_reduce(ref)
        for name in ('one_shot_all_reduce', 'two_shot_all_reduce_', 'multimem_all_reduce_', 'multimem_one_shot_all_reduce'):
            op = getattr(torch.ops.symm_mem, name, None)
            if op is None:
                res[name] = 'absent'
                continue
            try:
                saved = t.clone()
                o = op(t, 'sum', group.group_name)
                torch.cuda.synchronize()
                err = float((o - ref).abs().max() / ref.abs().max())
                t.copy_(saved)
                res[name + '_err'] = err
                res[name + '_eager_us'] = time_it(lambda: op(t, 'sum', group.group_name))
                t.copy_(saved)
                res[name + '_graph_us'] = graphed(lambda: op(t, 'sum', group.group_name))
                t.copy_(saved)
            except Exception as e:      # noqa: BLE001
                res[name] = repr(e)[:300]
    except Exception as e:      # noqa: BLE001
        res['symm_mem'] = repr(e)[:300]
    out[key] = res
    if rank == 0:
        print(key, json.dumps(res), flush=True)
if rank == 0:
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/allreduce_bench_N%d.json' % world, 'w') as f:
        json.dump(out, f, indent=1)
dist.barrier()
os._exit(0)

"""CPU, world_size 2 over gloo: the flat-bucket gradient exchange of stove_b200.dp
(averaging, dead parameters stay None, batch-sharding rule)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from stove_b200 import dp
    torch.manual_seed(rank)                       # different init per rank: broadcast must fix it
    model = torch.nn.ModuleDict({'a': torch.nn.Linear(5, 3), 'dead': torch.nn.Linear(2, 2)})
    eng = dp.DataParallel(model)
    full = torch.arange(40, dtype=torch.float32).view(8, 5) / 10
    mine = dp.shard(full)
    for p in eng.params:
        p.grad = None
    loss = (model['a'](mine) ** 2).mean()
    loss.backward()
    flat = eng.all_reduce_gradients()
    out[rank] = {'w': model['a'].weight.detach().clone(), 'b': model['a'].bias.detach().clone(),
                 'g': model['a'].weight.grad.clone(),
                 'dead': model['dead'].weight.grad is None, 'flat': flat.numel(), 'rows': mine.shape[0]}
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = out[0], out[1]
    assert torch.equal(r0['w'], r1['w'])                       # broadcast from rank 0
    assert torch.allclose(r0['g'], r1['g'])                    # same averaged gradient everywhere
    assert r0['dead'] and r1['dead'] and r0['flat'] == 18 and r0['rows'] == 4
    # equals the single-process gradient of the full (unsharded) batch
    lin = torch.nn.Linear(5, 3)
    with torch.no_grad():
        lin.weight.copy_(r0['w'])
        lin.bias.copy_(r0['b'])
    full = torch.arange(40, dtype=torch.float32).view(8, 5) / 10
    (lin(full) ** 2).mean().backward()
    assert torch.allclose(lin.weight.grad, r0['g'], atol=1e-6)

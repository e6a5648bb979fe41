"""GPU check + timing of the tensor-core LSTM kernels (csrc/lstm_tc.cu) through ops.LstmEncoder:
values / gradients against an fp64 LSTM, then per-kernel device times (library event profiler) of one
forward + backward at the config-1 shape.  Usage: python scripts/lstm_tc_bench.py [--skip-check]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stove_b200 import ops, _native as N


def ref(x, w_ih, w_hh, b_ih, b_hh, steps):
    n, H = x.shape[0], w_hh.shape[1]
    h = torch.zeros(n, H, dtype=x.dtype, device=x.device)
    c = torch.zeros_like(h)
    gx = x @ w_ih.t() + b_ih + b_hh
    outs = []
    for _ in range(steps):
        i, f, gg, o = (gx + h @ w_hh.t()).chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 1)


def params(K, H, dev='cuda'):
    return [((torch.rand(4 * H, K, device=dev) - 0.5) * 0.12).requires_grad_(True),
            ((torch.rand(4 * H, H, device=dev) - 0.5) * 0.12).requires_grad_(True),
            ((torch.rand(4 * H, device=dev) - 0.5) * 0.1).requires_grad_(True),
            ((torch.rand(4 * H, device=dev) - 0.5) * 0.1).requires_grad_(True)]


def check():
    torch.manual_seed(0)
    for M, Nn, K, parts in [(256, 128, 64, 1), (1024, 1024, 2048, 2), (2048, 256, 1024, 4), (100, 36, 40, 1)]:
        a, b = torch.randn(M, K, device='cuda'), torch.randn(Nn, K, device='cuda')
        d = ops.sum_parts(ops.tc3_gemm(ops.split_planes(a)[0], ops.split_planes(b)[0], parts=parts))
        r = a.double() @ b.double().t()
        print('tc3_gemm %dx%dx%d parts %d: rel err %.2e' % (M, Nn, K, parts, float((d.double() - r).abs().max() / r.abs().max())), flush=True)
    for n, K, H, steps in [(2048, 1024, 256, 3), (5, 1024, 256, 3), (300, 2500, 256, 3), (130, 1024, 64, 1), (2048, 1024, 256, 6)]:
        x = torch.rand(n, K, device='cuda')
        ps = params(K, H)
        out = ops.LstmEncoder.apply(x, *ps, steps)
        wgt = torch.randn_like(out)
        (out * wgt).sum().backward()
        pd = [p.detach().double().requires_grad_(True) for p in ps]
        o64 = ref(x.double(), *pd, steps)
        (o64 * wgt.double()).sum().backward()
        err = float((out.double() - o64).abs().max() / o64.abs().max())
        gerr = [float((a.grad.double() - b.grad).abs().max() / b.grad.abs().max()) for a, b in zip(ps, pd)]
        print('n=%d K=%d H=%d steps=%d: fwd rel err %.2e, grad rel err %s' % (n, K, H, steps, err, ['%.1e' % g for g in gerr]), flush=True)


def timing():
    n, K, H, steps = 2048, 1024, 256, 3
    x = torch.rand(n, K, device='cuda')
    ps = params(K, H)
    lib = N.lib()
    wgt = None
    for it in range(13):
        if it == 3:
            torch.cuda.synchronize()
            lib.stove_profile_enable(1)
            N.profile_read()
        out = ops.LstmEncoder.apply(x, *ps, steps)
        wgt = torch.randn_like(out) if wgt is None else wgt
        (out * wgt).sum().backward()
    torch.cuda.synchronize()
    lib.stove_profile_enable(0)
    recs = N.profile_read()
    per = collections.OrderedDict()
    order = []
    for name, ms in recs[:len(recs) // 10]:
        order.append(name)
    k = len(order)
    for i, name in enumerate(order):
        ts = sorted(recs[j * k + i][1] for j in range(10))
        print('%2d %-22s %7.1f us (median of 10, isolated launches)' % (i, name, 1e3 * ts[5]))
    print('sum %.1f us' % (1e3 * sum(sorted(recs[j * k + i][1] for j in range(10))[5] for i in range(k))))


if __name__ == '__main__':
    if '--skip-check' not in sys.argv:
        check()
    timing()

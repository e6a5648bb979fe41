"""GPU check of the tcgen05 LSTM step (csrc/lstm_tc.cu) through ops.LstmEncoder: forward and gradients
against an fp64 LSTM written out in torch, at several shapes, plus timing of the three forward launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stove_b200 import ops


def ref(x, w_ih, w_hh, b_ih, b_hh, steps):
    n, H = x.shape[0], w_hh.shape[1]
    h = torch.zeros(n, H, dtype=x.dtype, device=x.device)
    c = torch.zeros_like(h)
    gx = x @ w_ih.t() + b_ih + b_hh
    outs = []
    for _ in range(steps):
        g = gx + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 1)


def main():
    torch.manual_seed(0)
    dev = 'cuda'
    for n, K, H, steps in [(2048, 1024, 256, 3), (5, 1024, 256, 3), (300, 2500, 256, 3), (130, 1024, 64, 1), (2048, 1024, 256, 6)]:
        x = torch.rand(n, K, device=dev)
        w_ih = (torch.rand(4 * H, K, device=dev) - 0.5) * 0.12
        w_hh = (torch.rand(4 * H, H, device=dev) - 0.5) * 0.12
        b_ih = (torch.rand(4 * H, device=dev) - 0.5) * 0.1
        b_hh = (torch.rand(4 * H, device=dev) - 0.5) * 0.1
        ps = [w_ih, w_hh, b_ih, b_hh]
        for p in ps:
            p.requires_grad_(True)
        out = ops.LstmEncoder.apply(x, *ps, steps)
        wgt = torch.randn_like(out)
        (out * wgt).sum().backward()
        g = [p.grad.clone() for p in ps]
        pd = [p.detach().double().requires_grad_(True) for p in ps]
        o64 = ref(x.double(), *pd, steps)
        (o64 * wgt.double()).sum().backward()
        err = float((out.double() - o64).abs().max() / o64.abs().max())
        gerr = max(float((a.double() - b.grad).abs().max() / b.grad.abs().max()) for a, b in zip(g, pd))
        print('n=%d K=%d H=%d steps=%d: fwd rel err %.2e, grad rel err %.2e' % (n, K, H, steps, err, gerr), flush=True)
        # the same input GEMM through the library (cuBLAS TF32 over the same concatenated operands) and in SIMT fp32
        with torch.no_grad():
            xc, _ = ops.split_tf32_cat(x, 0, None)
            wc, _ = ops.split_tf32_cat(w_ih.detach(), 1, None)
            g_lib = ops._mm_tf32(xc, wc.t())
            g_f32 = x @ w_ih.detach().t()
            g64 = x.double() @ w_ih.detach().double().t()
            e_lib = float((g_lib.double() - g64).abs().max() / g64.abs().max())
            e_f32 = float((g_f32.double() - g64).abs().max() / g64.abs().max())
        print('   input GEMM alone: cuBLAS 3xTF32 rel err %.2e, cuBLAS fp32 %.2e' % (e_lib, e_f32), flush=True)
        assert err < 6e-5 and gerr < 2e-4, (err, gerr)
    # timing of the forward (3 launches + splits)
    n, K, H, steps = 2048, 1024, 256, 3
    x = torch.rand(n, K, device=dev)
    ps = [(torch.rand(4 * H, K, device=dev) - 0.5) * 0.1, (torch.rand(4 * H, H, device=dev) - 0.5) * 0.1,
          torch.zeros(4 * H, device=dev), torch.zeros(4 * H, device=dev)]
    with torch.no_grad():
        for _ in range(5):
            ops.LstmEncoder.apply(x, *ps, steps)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ops.LstmEncoder.apply(x, *ps, steps)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        print('encoder forward (graph replay): %.1f us' % (a.elapsed_time(b) * 20))
    from stove_b200 import _native as N
    lib = N.lib()
    lib.stove_profile_enable(1); N.profile_read()
    with torch.no_grad():
        for _ in range(10):
            ops.LstmEncoder.apply(x, *ps, steps)
    lib.stove_profile_enable(0)
    per = {}
    for name, t in N.profile_read():
        per.setdefault(name, []).append(t)
    for k, v in per.items():
        print('%-22s %3d launches, avg %.1f us, min %.1f us' % (k, len(v), 1e3 * sum(v) / len(v), 1e3 * min(v)))


if __name__ == '__main__':
    main()

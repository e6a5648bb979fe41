"""GPU parity of the recognition network kernels (encoder.py:28-57): the tcgen05 LSTM step with the cell
as its epilogue (csrc/lstm_tc.cu), the fused fc head (csrc/enc_head.cu) and the gradient-bucket gather,
against the same maths written out in fp64.  Tolerances: relative to the largest magnitude of the
reference tensor; 3xTF32 products carry ~1e-5 on dense inputs (the tensor cores truncate when aligning
addends), gradients 2e-4."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-300))


def _lstm_ref(x, w_ih, w_hh, b_ih, b_hh, steps):
    n, H = x.shape[0], w_hh.shape[1]
    h = torch.zeros(n, H, dtype=x.dtype, device=x.device)
    c = torch.zeros_like(h)
    gx = x @ w_ih.t() + b_ih + b_hh
    outs = []
    for _ in range(steps):
        i, f, g, o = (gx + h @ w_hh.t()).chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 1)


@pytest.mark.parametrize('n,K,H,steps', [(2048, 1024, 256, 3), (5, 1024, 256, 3), (300, 2500, 256, 6),
                                         (130, 1024, 64, 1), (129, 36, 32, 2)])
def test_lstm_encoder_vs_fp64(n, K, H, steps):
    from stove_b200 import ops
    torch.manual_seed(n + K)
    x = torch.rand(n, K, device='cuda') * (torch.rand(n, K, device='cuda') < 0.3)      # sparse frames, like the data
    ps = [((torch.rand(4 * H, K, device='cuda') - 0.5) * 0.12).requires_grad_(True),
          ((torch.rand(4 * H, H, device='cuda') - 0.5) * 0.12).requires_grad_(True),
          ((torch.rand(4 * H, device='cuda') - 0.5) * 0.1).requires_grad_(True),
          ((torch.rand(4 * H, device='cuda') - 0.5) * 0.1).requires_grad_(True)]
    out = ops.LstmEncoder.apply(x, *ps, steps)
    wgt = torch.randn_like(out)
    (out * wgt).sum().backward()
    pd = [p.detach().double().requires_grad_(True) for p in ps]
    ref = _lstm_ref(x.double(), *pd, steps)
    (ref * wgt.double()).sum().backward()
    assert out.shape == (n, steps, H)
    assert _rel(out, ref) < 3e-5
    for p, q in zip(ps, pd):
        assert _rel(p.grad, q.grad) < 2e-4


@pytest.mark.parametrize('M,Nn,K,parts', [(256, 128, 64, 1), (1024, 1024, 2048, 2), (2048, 256, 1024, 4), (100, 36, 40, 1),
                                          (1024, 2500, 300, 3)])
def test_tc3_gemm_vs_fp64(M, Nn, K, parts):
    """the 3xTF32 tcgen05 GEMM over (hi, lo) planes, split-K parts summed in a fixed order, and the transposed planes"""
    from stove_b200 import ops
    torch.manual_seed(M + K)
    a, b = torch.randn(M, K, device='cuda'), torch.randn(Nn, K, device='cuda')
    a_pl, aT_pl = ops.split_planes(a, True, True)
    assert torch.equal(a_pl[0] + a_pl[1], a) and torch.equal(a_pl[0], (a.view(torch.int32) & -8192).view(torch.float32))
    assert torch.equal(aT_pl[:, :, :M], a_pl.transpose(1, 2)) and not aT_pl[:, :, M:].any()
    b_pl = ops.split_planes(b)[0]
    d = ops.sum_parts(ops.tc3_gemm(a_pl, b_pl, parts=parts))
    ref = a.double() @ b.double().t()
    assert _rel(d, ref) < 1e-5
    for bn in (64, 256):                                # the other tile widths: same numbers up to summation order
        assert _rel(ops.sum_parts(ops.tc3_gemm(a_pl, b_pl, parts=parts, bn=bn)), ref) < 1e-5
    if K % 4 == 0 and M % 4 == 0:
        # contraction over the rows through the transposed planes (the weight-gradient form): a^T a
        dt = ops.sum_parts(ops.tc3_gemm(ops.split_planes(a.t().contiguous())[0], aT_pl[:, :, :], parts=1))
        assert _rel(dt, a.double().t() @ a.double()) < 5e-5      # all-positive diagonal sums: the accumulate bias of the tensor cores is coherent


@pytest.mark.parametrize('R,K,J,P', [(6144, 256, 50, 8), (7, 256, 50, 8), (100, 64, 20, 3), (33, 256, 64, 16)])
def test_enc_head_vs_fp64(R, K, J, P):
    from stove_b200 import ops
    torch.manual_seed(R)
    lead = (R // 3, 3) if R % 3 == 0 else (R,)                     # (frames, objects) or flat rows
    x = (torch.rand(*lead, K, device='cuda') - 0.5).requires_grad_(True)
    ps = [((torch.rand(J, K, device='cuda') - 0.5) * 0.3).requires_grad_(True),
          (torch.rand(J, device='cuda') - 0.5).requires_grad_(True),
          ((torch.rand(P, J, device='cuda') - 0.5) * 0.6).requires_grad_(True),
          (torch.rand(P, device='cuda') - 0.5).requires_grad_(True)]
    out = ops.EncHead.apply(x, *ps)
    wgt = torch.randn_like(out)
    (out * wgt).sum().backward()
    xd = x.detach().double().requires_grad_(True)
    pd = [p.detach().double().requires_grad_(True) for p in ps]
    ref = torch.sigmoid(xd @ pd[0].t() + pd[1]) @ pd[2].t() + pd[3]
    (ref * wgt.double()).sum().backward()
    assert out.shape == ref.shape
    assert _rel(out, ref) < 5e-6
    assert _rel(x.grad, xd.grad) < 1e-5
    for p, q in zip(ps, pd):
        assert _rel(p.grad, q.grad) < 2e-5


def test_enc_head_rejects_unsupported_shapes():
    from stove_b200 import ops
    x = torch.rand(4, 300, device='cuda')
    with pytest.raises(RuntimeError):
        ops.EncHead.apply(x, torch.rand(50, 300, device='cuda'), torch.rand(50, device='cuda'),
                          torch.rand(8, 50, device='cuda'), torch.rand(8, device='cuda'))


def test_gather_flat_equals_cat():
    from stove_b200 import ops
    torch.manual_seed(0)
    shapes = [(1,), (3, 5), (1024, 1024), (7,), (256, 1024), (50, 256), (0,), (33, 3)] + [(i + 1, 3) for i in range(140)]
    ts = [torch.rand(*s, device='cuda') for s in shapes]
    flat = ops.gather_flat(ts)
    assert torch.equal(flat, torch.cat([t.reshape(-1) for t in ts]))
    odd = [torch.rand(101, device='cuda')[1:], torch.rand(64, device='cuda')]        # unaligned source
    assert torch.equal(ops.gather_flat(odd), torch.cat(odd))


def test_recognition_net_single_node_vs_fp64():
    """LSTM + head as ONE autograd node (what RnnStates.forward uses): values and every gradient."""
    from stove_b200 import ops
    torch.manual_seed(3)
    n, K, H, steps, J, P = 700, 1024, 256, 3, 50, 8
    x = torch.rand(n, K, device='cuda') * (torch.rand(n, K, device='cuda') < 0.3)
    ps = [((torch.rand(4 * H, K, device='cuda') - 0.5) * 0.12), ((torch.rand(4 * H, H, device='cuda') - 0.5) * 0.12),
          ((torch.rand(4 * H, device='cuda') - 0.5) * 0.1), ((torch.rand(4 * H, device='cuda') - 0.5) * 0.1),
          ((torch.rand(J, H, device='cuda') - 0.5) * 0.3), (torch.rand(J, device='cuda') - 0.5),
          ((torch.rand(P, J, device='cuda') - 0.5) * 0.6), (torch.rand(P, device='cuda') - 0.5)]
    ps = [p.requires_grad_(True) for p in ps]
    out = ops.LstmEncoder.apply(x, *ps[:4], steps, *ps[4:])
    wgt = torch.randn_like(out)
    (out * wgt).sum().backward()
    pd = [p.detach().double().requires_grad_(True) for p in ps]
    h = _lstm_ref(x.double(), *pd[:4], steps)
    ref = torch.sigmoid(h @ pd[4].t() + pd[5]) @ pd[6].t() + pd[7]
    (ref * wgt.double()).sum().backward()
    assert out.shape == (n, steps, P)
    assert _rel(out, ref) < 3e-5
    for p, q in zip(ps, pd):
        assert _rel(p.grad, q.grad) < 2e-4

"""CPU: the hand-derived backward of the glimpse/mask kernel (tests/kernel_spec.py) against
autograd through the oracle."""
import math

import torch

from oracle import stove_oracle as so
import kernel_spec as ks

D = torch.float64


def test_scene_backward_formulas():
    g = torch.Generator().manual_seed(0)
    for align in (False, True):
        c = so.default_config(width=12, height=14, patch_width=5, patch_height=6, channels=2,
                              align_corners=align)
        F_, O = 3, 3
        img = torch.rand(F_, 2, 12, 14, generator=g, dtype=D)
        z = torch.zeros(F_, O, 4, dtype=D)
        z[..., 0] = 0.2 + 0.6 * torch.rand(F_, O, generator=g, dtype=D)
        z[..., 1] = 0.2 + 0.6 * torch.rand(F_, O, generator=g, dtype=D)
        z[..., 2:] = 0.9 * (2 * torch.rand(F_, O, 2, generator=g, dtype=D) - 1)
        z[0, 1] = z[0, 0] * 1.05
        zg = z.clone().requires_grad_(True)
        marg, bg, ov = so.masks_from_z(c, zg)
        patches = so.patches_from_z(c, img, zg.flatten(0, 1))
        wp = torch.rand(patches.shape, generator=g, dtype=D)
        wm = torch.rand(marg.shape, generator=g, dtype=D)
        wb = torch.rand(bg.shape, generator=g, dtype=D)
        wo = torch.rand(ov.shape, generator=g, dtype=D)
        ((patches * wp).sum() + (marg * wm).sum() + (bg * wb).sum() + (ov * wo).sum()).backward()
        for f in range(F_):
            p_, m_, b_, _, _ = ks.scene_frame_fwd(img[f], z[f], 5, 6, align)
            assert (p_ - patches.view(F_, O, 2, 5, 6)[f]).abs().max() < 1e-12
            assert (m_ - marg.view(F_, O, 2, 5, 6)[f, :, 0]).abs().max() < 1e-12
            assert (b_ - bg[f, 0]).abs().max() < 1e-12
            gm = wm.view(F_, O, 2, 5, 6)[f].sum(1) + wo[f][:, None, None] / 30.0
            gz = ks.scene_frame_bwd(img[f], z[f], 5, 6, wp.view(F_, O, 2, 5, 6)[f], gm,
                                    wb[f].sum(0), align)
            assert (gz - zg.grad[f]).abs().max() < 1e-9 * max(1.0, zg.grad[f].abs().max().item()), (align, f)


def test_fused_scene_likelihood_identities():
    """The algebra behind the fused kernels' shortcuts, against autograd in fp64: polynomial leaf table (forward value and
    the input pass), sums from the saved sum values, the chain rule of the sequence mode."""
    g = torch.Generator().manual_seed(1)
    G = 10
    mu = torch.rand(G, generator=g, dtype=D)
    var = 0.12 + 0.23 * torch.rand(G, generator=g, dtype=D)
    a, b = 0.5 / var, 0.5 * torch.log(var) + 0.9189385332046727
    c1, c2, c0 = ks.leaf_poly_table(mu, a, b)
    x = torch.rand((), generator=g, dtype=D).requires_grad_(True)
    m = torch.rand((), generator=g, dtype=D).requires_grad_(True)
    gl = torch.randn(G, generator=g, dtype=D)
    L = -(1 - m) * (a * (x - mu) ** 2 + b)
    assert torch.allclose(L, -(1 - m) * (c2 * x * x - c1 * x + c0), rtol=1e-12, atol=1e-12)
    (L * gl).sum().backward()
    gx, gm = ks.leaf_poly_backward(x.detach(), 1 - m.detach(), gl, c1, c2, c0)
    assert torch.allclose(gx, x.grad, rtol=1e-10) and torch.allclose(gm, m.grad, rtol=1e-10)

    l0, l1 = -30 * torch.rand(4, G, generator=g, dtype=D), -30 * torch.rand(4, G, generator=g, dtype=D)
    logw = torch.log_softmax(torch.randn(G * G, 3, generator=g, dtype=D), 0)
    prod = (l0.unsqueeze(1) + l1.unsqueeze(2)).reshape(4, -1)              # ProductVector: [j * G + i]
    sum_val = torch.logsumexp(prod.unsqueeze(-1) + logw, 1)
    e0, e1 = torch.exp(l0 - l0.max(-1, keepdim=True)[0]), torch.exp(l1 - l1.max(-1, keepdim=True)[0])
    T = torch.einsum('ni,nj,jis->ns', e0, e1, torch.exp(logw).view(G, G, 3))
    assert torch.allclose(ks.sums_from_saved(sum_val, l0, l1), T, rtol=1e-10)

    n, T_, skip, beta = 5, 8, 2, 10.0
    for t in (1, 4):
        z4 = torch.tensor([0.3, 1.1, -0.2, 0.4], dtype=D, requires_grad=True)
        out_obj = torch.tensor(-37.0, dtype=D, requires_grad=True)
        ov = torch.tensor(0.2, dtype=D, requires_grad=True)
        state = torch.stack([z4[0], z4[0] * z4[1], z4[2], z4[3]])       # sy_from_quotient, supair.py:151-158
        probe = torch.tensor([0.7, -1.3, 0.5, 2.0], dtype=D)             # stands in for the glimpse / mask dependence
        lik = out_obj * state[0] * state[1] + (math.log(beta) - beta * ov) + (probe * state).sum()
        elbo = lik / (n * (T_ - skip)) if t >= skip else lik / (n * (skip - 1))
        g_elbo = -1.0
        (g_elbo * elbo).backward()
        w = g_elbo / (n * (T_ - skip)) if t >= skip else g_elbo / (n * (skip - 1))
        fw, g_obj, g_ov, g_z = ks.sequence_mode_backward(g_elbo, n, T_, skip, t, z4.detach(), out_obj.detach(), w * probe, beta)
        assert abs(fw - w) < 1e-15 and torch.allclose(g_obj, out_obj.grad) and abs(g_ov - ov.grad) < 1e-12
        assert torch.allclose(g_z, z4.grad, rtol=1e-12)

"""GPU parity: glimpse/mask kernel and Supair.likelihood vs golden vectors and the oracle."""
import pytest
import torch

from oracle import stove_oracle as so
from util import Checker, load_golden, make_model

pytestmark = pytest.mark.gpu
VAL, GRAD = 2e-5, 2e-4


def test_scene_golden():
    g = load_golden('scene')
    oc, sd, model = make_model({}, int(g['seed']))
    x = so.bw_transform(g['x_u8'].double() / 255.0).float().cuda()
    z = g['z'].float().cuda().requires_grad_(True)
    ck = Checker('scene_golden')
    mp, mb, ov = model.sup.masks_from_z(z)
    ck.close('marg_patch', mp, g['marg_patch'], 1e-5, absolute=True)
    ck.close('marg_bg', mb, g['marg_bg'], 1e-5, absolute=True)
    ck.close('overlap', ov, g['overlap'], 1e-5, absolute=True)
    ck.close('patches', model.sup.patches_from_z(x.flatten(0, 1), z.flatten(0, 1)), g['patches'], 1e-5,
             absolute=True)
    model.sup.step_counter = 1
    ll, _ = model.sup.likelihood(x, z.flatten(0, 1))
    ck.close('ll', ll, g['ll'], VAL)
    model.zero_grad()
    (ll * g['w'].float().cuda()).sum().backward()
    ck.close('gz', z.grad, g['gz'], GRAD)
    params = dict(model.named_parameters())
    for k in g:
        if k.startswith('g.sup.'):
            ck.close(k, params[k[2:]].grad, g[k], GRAD)
    ck.finish()


@pytest.mark.parametrize('kw', [dict(align_corners=True), dict(width=50, height=50, num_obj=6),
                                dict(width=24, height=40, patch_width=6, patch_height=9, channels=3, num_obj=2)])
def test_scene_kernel_vs_oracle(kw):
    """align_corners=True (torch-1.0.1 semantics), 50x50 multiball and a non-square 3-channel case."""
    from stove_b200 import ops
    oc = so.default_config(**kw)
    gen = torch.Generator().manual_seed(1)
    F_, O, C = 7, oc.num_obj, oc.channels
    img = torch.rand(F_, C, oc.width, oc.height, generator=gen, dtype=torch.float64)
    z = torch.zeros(F_, O, 4, dtype=torch.float64)
    z[..., 0] = 0.1 + 0.7 * torch.rand(F_, O, generator=gen, dtype=torch.float64)
    z[..., 1] = 0.1 + 0.7 * torch.rand(F_, O, generator=gen, dtype=torch.float64)
    z[..., 2:] = 1.1 * (2 * torch.rand(F_, O, 2, generator=gen, dtype=torch.float64) - 1)
    zo = z.clone().requires_grad_(True)
    marg, bg, ov = so.masks_from_z(oc, zo)
    pat = so.patches_from_z(oc, img, zo.flatten(0, 1))
    ws = [torch.rand(t.shape, generator=gen, dtype=torch.float64) for t in (pat, marg, bg, ov)]
    sum((t * w).sum() for t, w in zip((pat, marg, bg, ov), ws)).backward()
    zg = z.float().cuda().requires_grad_(True)
    outs = ops.Scene.apply(img.float().cuda(), zg, oc.patch_width, oc.patch_height, oc.align_corners)
    sum((t * w.float().cuda()).sum() for t, w in zip(outs, ws)).backward()
    ck = Checker('scene_oracle_%s' % '_'.join(kw))
    for name, a, b in zip(('patches', 'marg_patch', 'marg_bg', 'overlap'), outs, (pat, marg, bg, ov)):
        ck.close(name, a, b, 2e-5, absolute=True)
    ck.close('gz', zg.grad, zo.grad, GRAD)
    ck.finish()


@pytest.mark.parametrize('align', [False, True])
def test_render_vs_grid_sample_and_oracle(align):
    """The renderer kernel (Supair.reconstruct_from_z's paste loop, supair.py:482-501) against the same loop written
    with F.affine_grid / F.grid_sample in fp64, per-frame and shared patches / backgrounds, boxes partly outside."""
    import torch.nn.functional as F
    from stove_b200 import ops
    torch.manual_seed(3)
    Fr, O, C, A, B, pa, pb = 37, 3, 1, 32, 32, 10, 10
    z = torch.cat([0.1 + 0.7 * torch.rand(Fr, O, 2), 2.2 * torch.rand(Fr, O, 2) - 1.1], -1).double()
    for per_frame in (False, True):
        bg = (torch.rand(Fr, C, A, B) if per_frame else torch.rand(C, A, B)).double() * 0.3
        patches = (torch.rand(Fr, O, C, pa, pb) if per_frame else torch.rand(O, C, pa, pb)).double()
        canvas = (bg if per_frame else bg.unsqueeze(0).repeat(Fr, 1, 1, 1)).clone()
        pf = patches if per_frame else patches.unsqueeze(0).repeat(Fr, 1, 1, 1, 1)
        for o in range(O):
            zz = z[:, o]
            inv = torch.stack([1 / zz[:, 0], 1 / zz[:, 1], -zz[:, 2] / zz[:, 0], -zz[:, 3] / zz[:, 1]], 1)
            zero = torch.zeros_like(inv[:, 0])
            theta = torch.stack([inv[:, 0], zero, inv[:, 2], zero, inv[:, 1], inv[:, 3]], 1).view(-1, 2, 3)
            grid = F.affine_grid(theta, torch.Size((Fr, C, A, B)), align_corners=align)
            canvas = canvas + F.grid_sample(pf[:, o], grid, align_corners=align)
        ref = canvas.clamp(0, 1)
        out = ops.render(bg.float().cuda(), patches.float().cuda(), z.float().cuda(), A, B, align)
        assert out.shape == ref.shape
        assert float((out.double().cpu() - ref).abs().max()) < 2e-5


def test_reconstruct_from_z_uses_the_renderer():
    """Supair.reconstruct_from_z on the GPU (renderer kernel) == the reference's grid_sample loop on the CPU copy."""
    from util import make_model
    oc, sd, model = make_model({}, 5)
    g = torch.Generator().manual_seed(1)
    z = torch.cat([0.15 + 0.5 * torch.rand(4, 6, 3, 2, generator=g), 1.6 * torch.rand(4, 6, 3, 2, generator=g) - 0.8], -1)
    gpu = model.sup.reconstruct_from_z(z.cuda())
    cpu = model.cpu().sup.reconstruct_from_z(z)
    assert gpu.shape == (4, 6, 1, 32, 32)
    assert float((gpu.cpu() - cpu).abs().max()) < 2e-5

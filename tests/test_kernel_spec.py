"""CPU: the hand-derived backward of the glimpse/mask kernel (tests/kernel_spec.py) against
autograd through the oracle."""
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

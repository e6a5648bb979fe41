"""GPU parity of the fused scene-likelihood kernels (csrc/scene_ll.cu, scene_ll_bwd.cu): one launch for
Supair.likelihood's op sequence (model/video_prediction/supair.py:62-76) against

  * the fp64 oracle (values and every gradient), and
  * the unfused launch sequence Scene -> Spn2 / Spn1 on the same inputs (every output of the three calls,
    including the activations saved for the backward pass).
"""
import pytest
import torch

from oracle import stove_oracle as so
from util import Checker, make_model

pytestmark = pytest.mark.gpu


def _inputs(oc, F_, seed, spread=1.1):
    gen = torch.Generator().manual_seed(seed)
    O = oc.num_obj
    # billiards-like sparse frames: mostly black with soft blobs (values in [0, 1]), some saturated
    img = torch.rand(F_, 1, oc.width, oc.height, generator=gen, dtype=torch.float64)
    img = (img * 2.5 - 1.4).clamp(0, 1)
    z = torch.zeros(F_, O, 4, dtype=torch.float64)
    z[..., 0] = 0.1 + (oc.max_obj_scale - 0.1) * torch.rand(F_, O, generator=gen, dtype=torch.float64)
    z[..., 1] = z[..., 0] * (0.75 + 0.5 * torch.rand(F_, O, generator=gen, dtype=torch.float64))
    z[..., 2:] = spread * (2 * torch.rand(F_, O, 2, generator=gen, dtype=torch.float64) - 1)
    w = [torch.rand(F_, generator=gen, dtype=torch.float64) + 0.5,
         torch.rand(F_ * O, generator=gen, dtype=torch.float64) + 0.5,
         torch.rand(F_, O, generator=gen, dtype=torch.float64) - 0.5]
    return img, z, w


def _run(model, img, z, w, fused):
    from stove_b200 import ops
    prev = ops.set_scene_ll(fused)
    try:
        model.zero_grad()
        zg = z.float().cuda().requires_grad_(True)
        bg, obj, ov, extra = model.sup.likelihood_raw(img.float().cuda(), zg)
        loss = (bg * w[0].float().cuda()).sum() + (obj * w[1].float().cuda()).sum() + (ov * w[2].float().cuda()).sum()
        loss.backward()
        grads = {k: p.grad.clone() for k, p in model.sup.named_parameters() if p.grad is not None}
        return dict(bg=bg.detach(), obj=obj.detach(), ov=ov.detach(), gz=zg.grad.clone(), grads=grads,
                    patches=extra['patches'].detach(), marg=extra['marginalise_flat'].detach(),
                    marg_bg=extra['marginalise_bg'].detach())
    finally:
        ops.set_scene_ll(prev)


@pytest.mark.parametrize('kw,F_', [({}, 1), ({}, 5), ({}, 33), ({}, 300), ({}, 1792),
                                   (dict(align_corners=True), 40),
                                   (dict(num_obj=6, width=50, height=50, max_obj_scale=0.22, debug_match_objects='greedy'), 70),
                                   (dict(num_obj=9, width=50, height=50, max_obj_scale=0.22, debug_match_objects='greedy'), 1792),
                                   (dict(num_obj=2, debug_match_objects='greedy'), 7)])
def test_fused_vs_unfused(kw, F_):
    """Every output and gradient of the fused path against the unfused kernels (both fp32: differences are
    summation-order only)."""
    from stove_b200 import ops
    oc, sd, model = make_model(kw, 31)
    img, z, w = _inputs(oc, F_, 5)
    a = _run(model, img, z, w, fused=True)
    b = _run(model, img, z, w, fused=False)
    ck = Checker('scene_ll_vs_unfused_%s_%d' % ('_'.join('%s%s' % kv for kv in kw.items()), F_))
    ck.true('fused path taken', ops.scene_ll_supported(img.float().cuda(), z.float().cuda(), oc.patch_width,
                                                       oc.patch_height, model.sup.obj_spn._tables.to(torch.device('cuda', 0)),
                                                       model.sup.bg_spn._tables.to(torch.device('cuda', 0))))
    for k in ('patches', 'marg', 'marg_bg', 'ov'):
        ck.close(k, a[k], b[k], 1e-4, absolute=True)      # both fp32; box edges amplify coordinate rounding by 1 / sx
    ck.close('bg', a['bg'], b['bg'], 3e-6)
    ck.close('obj', a['obj'], b['obj'], 3e-6)
    ck.mostly_close('gz', a['gz'], b['gz'], 2e-4)
    for k in b['grads']:
        ck.close('g.' + k, a['grads'][k], b['grads'][k], 2e-4)
    ck.finish()


@pytest.mark.parametrize('kw,F_', [({}, 29), (dict(num_obj=6, width=50, height=50, max_obj_scale=0.22, debug_match_objects='greedy'), 11)])
def test_fused_vs_oracle(kw, F_):
    oc, sd, model = make_model(kw, 32)
    img, z, w = _inputs(oc, F_, 6)
    a = _run(model, img, z, w, fused=True)
    from oracle.spn_oracle import spn_forward
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith('sup.') and 'output_vector' not in k}
    zo = z.clone().requires_grad_(True)
    obj_s, bg_s = so.structures(oc)
    marg_patch, marg_bg, ov = so.masks_from_z(oc, zo)                      # supair.py:62-76, term by term
    bg = spn_forward(bg_s, P, img.flatten(1), marg_bg.flatten(1), oc.bg_min_var, oc.bg_max_var, prefix='sup.bg_spn.')[:, 0]
    patches = so.patches_from_z(oc, img, zo.flatten(0, 1))
    obj = spn_forward(obj_s, P, patches.flatten(1), marg_patch.flatten(1), oc.obj_min_var, oc.obj_max_var,
                      prefix='sup.obj_spn.')[:, 0]
    ((bg * w[0]).sum() + (obj * w[1]).sum() + (ov * w[2]).sum()).backward()
    ck = Checker('scene_ll_vs_oracle_%s_%d' % ('_'.join(kw), F_))
    ck.close('bg', a['bg'], bg, 2e-5)
    ck.close('obj', a['obj'], obj, 2e-5)
    ck.close('ov', a['ov'], ov, 2e-5, absolute=True)
    ck.close('gz', a['gz'], zo.grad, 3e-4)
    for k, g in a['grads'].items():
        if 'sup.' + k in P and P['sup.' + k].grad is not None:
            ck.close('g.' + k, g, P['sup.' + k].grad, 3e-4)
    ck.finish()


def test_unsupported_configurations_take_the_unfused_kernels():
    """3-channel glimpses (object_embedding, stove.py:573-597) have no fused kernel: the query says so."""
    from stove_b200 import ops
    oc, sd, model = make_model({}, 31)
    dev = torch.device('cuda', 0)
    t2, t1 = model.sup.obj_spn._tables.to(dev), model.sup.bg_spn._tables.to(dev)
    img3 = torch.rand(4, 3, 32, 32, device='cuda')
    z = torch.rand(4, 3, 4, device='cuda')
    assert not ops.scene_ll_supported(img3, z, 10, 10, t2, t1)
    assert ops.scene_ll_supported(img3[:, :1].contiguous(), z, 10, 10, t2, t1)


@pytest.mark.parametrize('kw,n,T', [({}, 5, 8), ({}, 256, 8), (dict(num_obj=6, width=50, height=50, max_obj_scale=0.22,
                                                                       debug_match_objects='greedy'), 9, 5)])
def test_sequence_elbo_vs_composition(kw, n, T):
    """ops.SceneElbo (states read from z_sup / z_s, ELBO assembled in the kernel: one launch each way) against
    ZAll -> Scene -> Spn2 / Spn1 -> ElboAssemble (stove.py:731-748) on the same inputs: ELBO, the logging statistics
    and the gradients of z_sup, z_s, log q, the transition likelihood and every SPN parameter."""
    from stove_b200 import ops
    oc, sd, model = make_model(kw, 33)
    O, skip, Z = oc.num_obj, 2, oc.cl // 2 + 2
    S = T - skip
    img, z, _ = _inputs(oc, n * (T - 1), 8)
    gen = torch.Generator().manual_seed(9)
    zq = z.clone().view(n, T - 1, O, 4)
    zq[..., 1] = zq[..., 1] / zq[..., 0]                       # [sx, sy, x, y] -> [sx, sy / sx, x, y]
    z_sup = torch.cat([torch.rand(n, 1, O, 4, generator=gen, dtype=torch.float64) * 0.5 + 0.1, zq], 1)
    z_sup[:, skip:] = torch.rand(n, S, O, 4, generator=gen, dtype=torch.float64) * 0.5 + 0.1      # unused rows
    z_s = torch.cat([zq[:, skip - 1:], torch.randn(n, S, O, Z - 4, generator=gen, dtype=torch.float64)], -1)
    logq = torch.randn(n, S, generator=gen, dtype=torch.float64) * 3
    trans = torch.randn(n, S, generator=gen, dtype=torch.float64) * 3
    x_img = img.float().cuda()
    packed_done = []

    def run(fused):
        model.zero_grad()
        leaves = [t.float().cuda().requires_grad_(True) for t in (z_sup, z_s, logq, trans)]
        zs_, zz_, lq_, tr_ = leaves
        packed = model.sup.pack()
        packed_done.append(packed)
        if fused:
            out = model.sup.sequence_elbo(x_img, zs_, zz_, lq_, tr_, skip, packed)
            assert out is not None, 'fused sequence path not taken'
            elbo, stats = out[0], out[1]
        else:
            prev = ops.set_scene_ll(False)
            try:
                z_all = ops.ZAll.apply(zs_, zz_, skip)
                bg, obj, ov, _ = model.sup.likelihood_raw(x_img, z_all.flatten(end_dim=1), packed=packed)
                elbo, stats = ops.ElboAssemble.apply(bg, obj, z_all, ov, lq_, tr_, skip, float(oc.overlap_beta))
            finally:
                ops.set_scene_ll(prev)
        (-elbo).backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.clone() for k, p in model.sup.named_parameters() if p.grad is not None}
        return elbo.detach(), stats.detach(), [t.grad.clone() for t in leaves], grads

    ea, sa, ga, pa_ = run(True)
    eb, sb, gb, pb_ = run(False)
    ck = Checker('scene_elbo_vs_composition_%s_%d' % ('_'.join(kw), n))
    ck.close('elbo', ea, eb, 2e-6)
    ck.close('stats', sa[:7], sb[:7], 2e-6)
    for name, a, b in zip(('g_z_sup', 'g_z_s', 'g_logq', 'g_trans'), ga, gb):
        ck.mostly_close(name, a, b, 2e-4)
    for k in pb_:
        ck.close('g.' + k, pa_[k], pb_[k], 2e-4)
    ck.finish()


def test_fused_forward_with_unfused_backward():
    """The fused forward kernel leaves the saved activations in the layouts of spn_obj.cu / spn_bg.cu: differentiating it
    with the UNFUSED backward kernels (ops.set_scene_ll_bwd(False)) gives the gradients of the fused backward kernel."""
    from stove_b200 import ops
    oc, sd, model = make_model({}, 34)
    img, z, w = _inputs(oc, 45, 11)
    a = _run(model, img, z, w, fused=True)
    prev = ops.set_scene_ll_bwd(False)
    try:
        b = _run(model, img, z, w, fused=True)
    finally:
        ops.set_scene_ll_bwd(prev)
    ck = Checker('scene_ll_fused_fwd_unfused_bwd')
    ck.close('bg', a['bg'], b['bg'], 1e-7)
    ck.close('obj', a['obj'], b['obj'], 1e-7)
    ck.mostly_close('gz', a['gz'], b['gz'], 2e-4)
    for k in b['grads']:
        ck.close('g.' + k, a['grads'][k], b['grads'][k], 2e-4)
    ck.finish()

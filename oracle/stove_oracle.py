"""ORACLE (test infrastructure only) -- CPU restatement of STOVE's hot path.

This file is the *checker*, never the product: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg may import it.

Parity status: **pinned** against the reference itself run in the build container
(`oracle/make_golden.py` -> `tests/golden/*.npz`; `tests/test_oracle_vs_reference.py`
re-checks live when `/root/reference` is mounted).  The reference has no tests of its own.

The model is a plain function of a parameter dict that uses the reference's
`state_dict()` names, so the same weights drive the reference, this oracle and the CUDA
modules.  Restated (reference file:line):
  * bw_transform                 model/utils/utils.py:10-15
  * RnnStates encoder            model/video_prediction/encoder.py:28-57
  * Supair.constrain_zp          model/video_prediction/supair.py:112-149
  * patches_from_z / expand_z    model/video_prediction/supair.py:193-216, 241-276
  * masks_from_z / invert_z      model/video_prediction/supair.py:218-239, 278-356
  * Supair.likelihood            model/video_prediction/supair.py:44-110
  * Dynamics.forward / core      model/video_prediction/dynamics.py:181-265
  * Dynamics.constrain_z_dyn     model/video_prediction/dynamics.py:147-179
  * match / fix / velocities     model/video_prediction/stove.py:54-101, 200-571
  * full_state, transition_lik   model/video_prediction/stove.py:103-198
  * stove_forward (ELBO)         model/video_prediction/stove.py:599-775
  * rollout                      model/video_prediction/stove.py:777-861
`F.affine_grid` + `F.grid_sample` (third-party, torch) are restated in closed form
(`affine_bilinear`), with the `align_corners` switch of SURVEY.md hard part 1.
"""
import math
from types import SimpleNamespace

import torch

from .spn_oracle import obj_spn_structure, bg_spn_structure, spn_forward

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# configuration (defaults of model/video_prediction/config.py:6-134)
# --------------------------------------------------------------------------------------
def default_config(**kw):
    c = SimpleNamespace(
        channels=1, cl=32, num_obj=3, width=32, height=32, skip=2,
        patch_height=10, patch_width=10,
        obj_min_var=0.12, obj_max_var=0.35, bg_min_var=0.002, bg_max_var=0.16,
        scale_var=0.3, pos_var=0.3, min_obj_scale=0.1, max_obj_scale=0.8,
        min_y_scale=0.75, max_y_scale=1.25, obj_pos_bound=0.9,
        obj_spn_num_gauss=10, obj_spn_num_sums=10, overlap_beta=10.0,
        transition_lik_std=[0.01, 0.01, 0.01, 0.01], debug_latent_q_std=0.04,
        debug_fix_supair=True, debug_match_appearance=False, debug_core_appearance=False,
        debug_appearance_dim=3, debug_match_objects='3_only', debug_bw=True,
        debug_nonlinear='relu', debug_no_latents=False, debug_no_reuse=False,
        debug_no_velocity=False, action_conditioned=False, action_space=None,
        random_seed=7, num_rollout=8, align_corners=False)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def structures(c):
    obj = obj_spn_structure(c.channels * c.patch_width * c.patch_height, c.random_seed,
                            c.obj_spn_num_gauss, c.obj_spn_num_sums)
    bg = bg_spn_structure(c.width * c.height * c.channels, c.random_seed)
    return obj, bg


# --------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------
def bw_transform(x):
    """utils.py:10-15: sum colour channels, clamp to [0, 1], keep a channel dim."""
    return torch.clamp(x.sum(2), 0, 1).unsqueeze(2)


def normal_log_prob(value, mean, std):
    return -((value - mean) ** 2) / (2 * std ** 2) - torch.log(std) - _HALF_LOG_2PI


def affine_bilinear(img, z, out_a, out_b, align_corners=False):
    """`F.grid_sample(img, F.affine_grid([[sx,0,x],[0,sy,y]], (N,C,out_a,out_b)))`
    with bilinear interpolation and zero padding (supair.py:272-275), in closed form.

    img (N, C, A, B); z (N, 4) = [sx, sy, x, y].  The x coordinate walks the LAST image
    dim (B), y the second-to-last (A) -- the reference ignores the w/h naming
    (supair.py:244-247).
    """
    N, C, A, B = img.shape
    dt = img.dtype
    j = torch.arange(out_b, dtype=dt)
    i = torch.arange(out_a, dtype=dt)
    if align_corners:
        xb = 2 * j / max(out_b - 1, 1) - 1
        yb = 2 * i / max(out_a - 1, 1) - 1
    else:
        xb = (2 * j + 1) / out_b - 1
        yb = (2 * i + 1) / out_a - 1
    gx = z[:, 0:1] * xb[None, :] + z[:, 2:3]                 # (N, out_b)
    gy = z[:, 1:2] * yb[None, :] + z[:, 3:4]                 # (N, out_a)
    if align_corners:
        px = (gx + 1) * (B - 1) / 2
        py = (gy + 1) * (A - 1) / 2
    else:
        px = ((gx + 1) * B - 1) / 2
        py = ((gy + 1) * A - 1) / 2
    x0 = torch.floor(px)
    y0 = torch.floor(py)
    fx = px - x0
    fy = py - y0
    out = img.new_zeros(N, C, out_a, out_b)
    flat = img.reshape(N, C, A * B)
    for dy, wy in ((0, 1 - fy), (1, fy)):
        for dx, wx in ((0, 1 - fx), (1, fx)):
            xi = x0 + dx
            yi = y0 + dy
            okx = (xi >= 0) & (xi <= B - 1)
            oky = (yi >= 0) & (yi <= A - 1)
            xi = xi.clamp(0, B - 1).long()
            yi = yi.clamp(0, A - 1).long()
            idx = (yi[:, :, None] * B + xi[:, None, :]).reshape(N, 1, -1).expand(N, C, -1)
            val = torch.gather(flat, 2, idx).reshape(N, C, out_a, out_b)
            w = (wy * oky.to(dt))[:, :, None] * (wx * okx.to(dt))[:, None, :]
            out = out + val * w[:, None]
    return out


def invert_z(z):
    """supair.py:218-239."""
    return torch.stack([1 / z[:, 0], 1 / z[:, 1], -z[:, 2] / z[:, 0], -z[:, 3] / z[:, 1]], 1)


def sy_from_quotient(z):
    """supair.py:151-156."""
    return torch.cat([z[..., 0:1], z[..., 0:1] * z[..., 1:2], z[..., 2:]], -1)


def constrain_zp(c, zp):
    """supair.py:112-149."""
    mean = torch.cat([torch.sigmoid(zp[:, 0:2]), 2 * torch.sigmoid(zp[:, 2:4]) - 1], -1)
    std = torch.cat([c.scale_var * torch.sigmoid(zp[:, 4:6]),
                     c.pos_var * torch.sigmoid(zp[:, 6:8])], -1)
    hi = zp.new_tensor([[c.max_obj_scale - c.min_obj_scale, c.max_y_scale - c.min_y_scale,
                         c.obj_pos_bound, c.obj_pos_bound]])
    lo = zp.new_tensor([[c.min_obj_scale, c.min_y_scale, 0.0, 0.0]])
    return mean * hi + lo, std


# --------------------------------------------------------------------------------------
# scene model
# --------------------------------------------------------------------------------------
def patches_from_z(c, x_img, z_obj):
    """supair.py:241-276.  x_img (F, C, A, B), z_obj (F*O, 4) -> (F*O, C, pw, ph)."""
    O = z_obj.shape[0] // x_img.shape[0]
    x_rep = x_img.repeat_interleave(O, 0)
    return affine_bilinear(x_rep, z_obj, c.patch_width, c.patch_height, c.align_corners)


def masks_from_z(c, z_img):
    """supair.py:278-356.  z_img (F, O, 4) -> marg_patch (F*O,C,pw,ph), bg (F,C,A,B),
    overlap (F, O)."""
    F_, O = z_img.shape[0], z_img.shape[1]
    ones = z_img.new_ones(F_, c.channels, c.width, c.height)
    bg = z_img.new_zeros(F_, c.channels, c.width, c.height)
    margs = []
    for o in range(O):
        zo = z_img[:, o]
        margs.append(1.0 - affine_bilinear(1.0 - bg, zo, c.patch_width, c.patch_height,
                                           c.align_corners))
        paste = affine_bilinear(ones, invert_z(zo), c.width, c.height, c.align_corners)
        bg = torch.clamp(bg + paste, 0, 1)
    marg = torch.stack(margs, 1)
    overlap = marg.flatten(2).mean(2)
    return marg.flatten(0, 1), bg, overlap


def likelihood(c, P, structs, x, z_obj, parts=None):
    """supair.py:44-110.  x (n, T', C, A, B), z_obj (n*T'*O, 4) [sx, sy, x, y]."""
    obj_s, bg_s = structs
    x_img = x.flatten(0, 1)
    z_img = z_obj.view(-1, c.num_obj, 4)
    marg_patch, marg_bg, overlap = masks_from_z(c, z_img)
    bg_ll = spn_forward(bg_s, P, x_img.flatten(1), marg_bg.flatten(1),
                        c.bg_min_var, c.bg_max_var, prefix='sup.bg_spn.')[:, 0]
    patches = patches_from_z(c, x_img, z_obj)
    p_ll = spn_forward(obj_s, P, patches.flatten(1), marg_patch.flatten(1),
                       c.obj_min_var, c.obj_max_var, prefix='sup.obj_spn.')[:, 0]
    p_ll = (p_ll * z_obj[:, 0] * z_obj[:, 1]).view(-1, c.num_obj).sum(1)
    beta = c.overlap_beta
    ov_ll = (math.log(beta) - beta * overlap).sum(1)           # Exponential(beta).log_prob
    if parts is not None:
        parts.update(bg=bg_ll, patch=p_ll, overlap=ov_ll, patches=patches,
                     marg_patch=marg_patch, marg_bg=marg_bg, overlap_ratios=overlap)
    return bg_ll + p_ll + ov_ll


# --------------------------------------------------------------------------------------
# encoder (nn.LSTM restated; gate order i, f, g, o)
# --------------------------------------------------------------------------------------
def encoder(c, P, frames):
    """encoder.py:28-57: the same flattened frame is fed for num_obj LSTM steps."""
    x = frames.flatten(1)
    Wih, Whh = P['sup.encoder.rnn.weight_ih_l0'], P['sup.encoder.rnn.weight_hh_l0']
    b = P['sup.encoder.rnn.bias_ih_l0'] + P['sup.encoder.rnn.bias_hh_l0']
    H = Whh.shape[1]
    h = x.new_zeros(x.shape[0], H)
    cell = x.new_zeros(x.shape[0], H)
    xin = x @ Wih.t() + b
    outs = []
    for _ in range(c.num_obj):
        g = xin + h @ Whh.t()
        i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        cell = torch.sigmoid(f) * cell + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(cell)
        outs.append(h)
    out = torch.stack(outs, 1)
    out = torch.sigmoid(out @ P['sup.encoder.fc1.weight'].t() + P['sup.encoder.fc1.bias'])
    return out @ P['sup.encoder.fc2.weight'].t() + P['sup.encoder.fc2.bias']


# --------------------------------------------------------------------------------------
# dynamics
# --------------------------------------------------------------------------------------
def _lin(P, name, v):
    return v @ P[name + '.weight'].t() + P[name + '.bias']


def _phi(c, v):
    # dynamics.py:109 -- the selector is inverted: everything except 'leaky_relu' gives
    # F.leaky_relu (slope 0.01); 'leaky_relu' gives F.elu.
    if c.debug_nonlinear == 'leaky_relu':
        return torch.nn.functional.elu(v)
    return torch.where(v >= 0, v, 0.01 * v)


def dynamics_forward(c, P, s, core_idx=0, actions=None, obj_appearances=None, lim_enc=2,
                     parts=None):
    """dynamics.py:220-265 (+ core :181-218).  s (n, O, cl//2) -> (n, O, cl), reward."""
    O = c.num_obj
    k = 'dyn.'
    if actions is not None:
        emb = _lin(P, k + 'action_embedding_layer', actions).view(actions.shape[0], O, 4)
        s = torch.cat([s, emb], -1)
    if obj_appearances is not None:
        s = torch.cat([s, obj_appearances], -1)
    s = torch.cat([s[..., :lim_enc], _lin(P, k + 'state_enc', s)[..., lim_enc:]], -1)

    ci = str(core_idx)
    h = _phi(c, _lin(P, k + 'self_cores.%s.0' % ci, s))
    self_dyn = _lin(P, k + 'self_cores.%s.1' % ci, h) + h
    a1 = s.unsqueeze(2).expand(-1, -1, O, -1)        # [b, i, j] = s_i
    a2 = s.unsqueeze(1).expand(-1, O, -1, -1)        # [b, i, j] = s_j
    dist = ((a1[..., 0] - a2[..., 0]) ** 2 + (a1[..., 1] - a2[..., 1]) ** 2).unsqueeze(-1)
    comb = torch.cat([a1, a2, dist], 3)
    r = _phi(c, _lin(P, k + 'rel_cores.%s.0' % ci, comb))
    r = _phi(c, _lin(P, k + 'rel_cores.%s.1' % ci, r))
    rel = _lin(P, k + 'rel_cores.%s.2' % ci, r) + r
    a = _phi(c, _lin(P, k + 'att_net.%s.0' % ci, comb))
    a = _phi(c, _lin(P, k + 'att_net.%s.1' % ci, a))
    att = torch.exp(_lin(P, k + 'att_net.%s.2' % ci, a))
    mask = (1 - torch.eye(O, dtype=s.dtype)).view(1, O, O, 1)
    rel_dyn = (rel * mask * att).sum(2)
    d = self_dyn + rel_dyn
    f1 = torch.tanh(_lin(P, k + 'affector.%s.0' % ci, d))
    f2 = torch.tanh(_lin(P, k + 'affector.%s.1' % ci, f1)) + f1
    f3 = _lin(P, k + 'affector.%s.2' % ci, f2)
    o1 = torch.tanh(_lin(P, k + 'out.%s.0' % ci, torch.cat([f3, s], 2)))
    result = _lin(P, k + 'out.%s.1' % ci, o1) + o1
    if parts is not None:
        parts.update(s_enc=s, dynamic_pred=d)
    if c.action_conditioned:
        r0 = _lin(P, k + 'reward_head0.2', torch.relu(_lin(P, k + 'reward_head0.0', d))).sum(1)
        r1 = torch.relu(_lin(P, k + 'reward_head1.0', r0))
        r1 = torch.relu(_lin(P, k + 'reward_head1.2', r1))
        reward = torch.sigmoid(_lin(P, k + 'reward_head1.4', r1).view(-1, 1))
        return result, reward
    return result, 0


def constrain_z_dyn(c, z, z_std=None):
    """dynamics.py:147-179."""
    zc = 2 * torch.sigmoid(z) - 1
    if z_std is None:
        return zc, None
    std = torch.cat([c.pos_var * torch.sigmoid(z_std[..., :2]),
                     0.04 * torch.sigmoid(z_std[..., 2:4]),
                     c.debug_latent_q_std * torch.sigmoid(z_std[..., 4:])], -1)
    return zc, std


def transition_std(c, like):
    std = list(c.transition_lik_std)
    if len(std) == 4:
        std = std + 12 * [0.01]                                     # dynamics.py:112-120
    return like.new_tensor([[std]])


# --------------------------------------------------------------------------------------
# sequence glue of stove.py
# --------------------------------------------------------------------------------------
def _match_prepare(c, z_sup, z_sup_std, app):
    z = (z_sup + 1) / 2
    m_idx = [2, 3]
    if app is not None:
        z = torch.cat([z, app], -1)
        if c.debug_match_appearance:
            m_idx += [4, 5, 6]
    if z_sup_std is not None:
        z = torch.cat([z, z_sup_std], -1)
    return z, m_idx


def _match_finish(z_matched, z_sup_std, app):
    z_sup_m = 2 * z_matched[..., :4] - 1
    app_m = z_matched[..., 4:7] if app is not None else None
    if z_sup_std is None and app is not None:
        return z_sup_m, app_m
    if z_sup_std is not None and app is None:
        return z_sup_m, z_matched[..., 4:8], None
    if z_sup_std is not None and app is not None:
        return z_sup_m, z_matched[..., 7:11], app_m
    return z_sup_m


def match_3_only(c, z_sup, z_sup_std=None, app=None):
    """stove.py:200-329: per-step nearest neighbour, greedy repair of non-permutations."""
    z, m_idx = _match_prepare(c, z_sup, z_sup_std, app)
    O = c.num_obj
    out = [z[:, 0]]
    for t in range(1, z.shape[1]):
        curr = z[:, t][..., m_idx].detach().unsqueeze(1)            # [b, 1, cur, :]
        prev = out[t - 1][..., m_idx].detach().unsqueeze(2)         # [b, prev, 1, :]
        err = ((prev - curr) ** 2).sum(-1)                          # [b, prev, cur]
        idx = err.min(-1)[1]
        ok = (idx[:, 0] != idx[:, 1]) & (idx[:, 1] != idx[:, 2]) & (idx[:, 0] != idx[:, 2])
        bad = (~ok).nonzero().flatten()
        if bad.numel() > 0:
            e = err[bad].clone()
            fixed = torch.zeros(bad.numel(), O, dtype=torch.long)
            for o in range(O):
                # stove.py:278-295: row o takes its current argmin; that column is then
                # knocked out for every row
                col = e.min(-1)[1][:, o]
                fixed[:, o] = col
                e[torch.arange(bad.numel()), :, col] = 1e12
            idx[bad] = fixed
        out.append(torch.gather(z[:, t], 1, idx.unsqueeze(-1).expand(-1, -1, z.shape[-1])))
    return _match_finish(torch.stack(out, 1), z_sup_std, app)


def match_greedy(c, z_sup, z_sup_std=None, app=None):
    """stove.py:432-514: global greedy bipartite matching."""
    z, m_idx = _match_prepare(c, z_sup, z_sup_std, app)
    O = c.num_obj
    zm = torch.zeros_like(z)
    zm[:, 0] = z[:, 0]
    for t in range(1, z.shape[1]):
        curr = z[:, t][..., m_idx].detach().unsqueeze(1)
        prev = zm[:, t - 1][..., m_idx].detach().unsqueeze(2)
        err = ((prev - curr) ** 2).sum(-1).clone()
        n = err.shape[0]
        perm = torch.zeros_like(err)
        rows = torch.arange(n)
        for _ in range(O):
            flat = err.view(n, -1).argmin(1)
            ix, iy = flat // O, flat % O
            perm[rows, ix, iy] = 1.0
            err[rows, ix, :] = err.max() + 1
            err[rows, :, iy] = err.max() + 1
        zm[:, t] = perm @ z[:, t]
    return _match_finish(zm, z_sup_std, app)


def match_volatile(c, z_sup, z_sup_std=None, app=None):
    """stove.py:331-430: per-object nearest neighbour without permutation check."""
    z, m_idx = _match_prepare(c, z_sup, z_sup_std, app)
    O = c.num_obj
    out = [z[:, 0]]
    for t in range(1, z.shape[1]):
        curr = z[:, t][..., m_idx].detach().unsqueeze(2)            # [b, cur, 1, :]
        prev = out[t - 1][..., m_idx].detach().unsqueeze(1)         # [b, 1, prev, :]
        err = ((prev - curr) ** 2).sum(-1)                          # [b, cur, prev]
        col = err.min(-2)[1]                                        # per prev: best cur
        perm = torch.zeros(z.shape[0] * O * O, dtype=z.dtype)
        perm[torch.arange(col.numel()) * O + col.flatten()] = 1
        out.append(perm.view(-1, O, O) @ z[:, t])
    return _match_finish(torch.stack(out, 1), z_sup_std, app)


MATCHERS = {'3_only': match_3_only, 'greedy': match_greedy, 'volatile': match_volatile}


def fix_supair(z, z_std):
    """stove.py:516-571."""
    zz = torch.cat([z, z_std], -1).clone()
    d = (zz[:, 1:, :, :2] - zz[:, :-1, :, :2]).abs().detach()
    zero = torch.zeros_like(d[:, :1])
    flag = (torch.cat([zero, d], 1) > 0.095) & (torch.cat([d, zero], 1) > 0.095)
    smooth = (zz[:, :-2] + zz[:, 2:]) / 2
    pad = torch.zeros_like(zz[:, :1])
    smooth = torch.cat([pad, smooth, pad], 1)
    flag = torch.cat(zz.shape[-1] // 2 * [flag], -1)
    zz = torch.where(flag, smooth, zz)
    return torch.chunk(zz, 2, dim=-1)


def v_from_state(z_sup):
    """stove.py:54-80."""
    v = z_sup[:, 1:, :, 2:] - z_sup[:, :-1, :, 2:]
    full = torch.cat([z_sup[:, 1:], v], -1)
    return torch.cat([torch.zeros_like(full[:, :1]), full], 1)


def v_std_from_pos(z_sup_std):
    """stove.py:82-101."""
    v = torch.sqrt(z_sup_std[:, 1:, :, 2:] ** 2 + z_sup_std[:, :-1, :, 2:] ** 2)
    full = torch.cat([z_sup_std[:, 1:], v], -1)
    return torch.cat([torch.zeros_like(full[:, :1]), full], 1)


def full_state(c, z_dyn, std_dyn, z_sup, std_sup, eps):
    """stove.py:103-170 (default branch: latents sampled, velocities kept)."""
    m_sup, s_sup = z_sup[..., 2:6], std_sup[..., 2:6]
    m_dyn, s_dyn = z_dyn[..., :4], std_dyn[..., :4]
    mean_xv = (s_sup ** 2 * m_dyn + s_dyn ** 2 * m_sup) / (s_dyn ** 2 + s_sup ** 2)
    std_xv = s_dyn * s_sup / torch.sqrt(s_dyn ** 2 + s_sup ** 2)
    mean = torch.cat([z_sup[..., :2], mean_xv, z_dyn[..., 4:]], -1)
    std = torch.cat([std_sup[..., :2], std_xv, std_dyn[..., 4:]], -1)
    z = mean + std * eps
    return z, normal_log_prob(z, mean, std), mean, std


def object_embedding(c, z, x_color):
    """stove.py:573-597: mean colour of each object's glimpse."""
    zp = sy_from_quotient(z[..., :4].detach())
    cc = SimpleNamespace(**vars(c))
    patches = patches_from_z(cc, x_color.flatten(0, 1), zp.flatten(0, 2))
    return patches.mean((-1, -2)).view(*z.shape[:-1], 3)


def stove_forward(c, P, x, noise, actions=None, structs=None, parts=None):
    """Stove.forward + stove_forward, stove.py:863-897 and 599-775.

    noise: list of standard-normal draws in the reference's order:
      [ (n,O,12,1) latent prior, (n,O,12,1) std prior, (n,O,18) per t in skip..T-1 ].
    Returns elbo (0-dim), dict of latents, rewards.
    """
    structs = structs or structures(c)
    x_color = x
    if c.debug_bw:
        x = bw_transform(x)
    n, T = x.shape[0], x.shape[1]
    skip, cl, O = c.skip, c.cl, c.num_obj

    zp = encoder(c, P, x.flatten(0, 1))
    z_sup, z_sup_std = constrain_zp(c, zp.flatten(0, 1))
    z_sup = z_sup.view(n, T, O, 4)
    z_sup_std = z_sup_std.view(n, T, O, 4)
    app = None
    if c.debug_core_appearance or c.debug_match_appearance:
        app = object_embedding(c, z_sup, x_color)
    z_sup, z_sup_std, app = MATCHERS[c.debug_match_objects](c, z_sup, z_sup_std, app)
    core_app = app.transpose(0, 1) if c.debug_core_appearance else T * [None]
    if c.debug_fix_supair:
        z_sup, z_sup_std = fix_supair(z_sup, z_sup_std)
    z_sup_full = v_from_state(z_sup)
    z_sup_std_full = v_std_from_pos(z_sup_std)

    noise = list(noise)
    lat0 = (0.0 + 0.01 * noise.pop(0)).squeeze()
    init_z = torch.cat([z_sup_full[:, skip - 1], lat0], -1)
    std0 = (0.1 + 0.01 * noise.pop(0)).squeeze()
    dyn_std_init = torch.cat([z_sup_std_full[:, skip - 1, :, 2:], std0], -1)

    z = {skip - 1: init_z}
    z_dyn, z_dyn_std, z_std, log_z, rewards = {}, {skip - 1: dyn_std_init}, {}, {}, []
    core_act = actions.transpose(0, 1) if actions is not None else T * [None]
    for t in range(skip, T):
        tmp, reward = dynamics_forward(c, P, z[t - 1][..., 2:], 0, core_act[t - 1], core_app[t - 1])
        rewards.append(reward)
        zd, z_dyn_std[t] = constrain_z_dyn(c, tmp[..., :cl // 2], tmp[..., cl // 2:])
        z_dyn[t] = torch.cat([z[t - 1][..., 2:4] + zd[..., :2], zd[..., 2:]], -1)
        z[t], log_z[t], _, z_std[t] = full_state(
            c, z_dyn[t], z_dyn_std[t], z_sup_full[:, t], z_sup_std_full[:, t], noise.pop(0))
    rng = range(skip, T)
    z_s = torch.stack([z[t] for t in rng], 1)
    z_dyn_s = torch.stack([z_dyn[t] for t in rng], 1)
    log_z_s = torch.stack([log_z[t] for t in rng], 1)
    if c.action_conditioned:
        rewards = torch.stack(rewards, 1)
    else:
        rewards = torch.zeros(len(rewards))

    z_f = sy_from_quotient(z_s.flatten(0, 2))
    lik_parts = {} if parts is not None else None
    img_lik = likelihood(c, P, structs, x[:, skip:], z_f[..., :4], lik_parts)
    z_sup_tmp = sy_from_quotient(z_sup[:, 1:skip])
    img_lik_sup = likelihood(c, P, structs, x[:, 1:skip], z_sup_tmp.flatten(0, 2))
    log_z_f = log_z_s.sum((-2, -1)).flatten()
    trans = normal_log_prob(z_s[..., 2:], z_dyn_s, transition_std(c, z_s))
    trans = trans.sum((-2, -1)).flatten(0, 1)
    elbo = trans + img_lik - log_z_f
    average_elbo = elbo.mean() + img_lik_sup.mean()
    prop = {
        'z': sy_from_quotient(z_s).detach(),
        'z_dyn': z_dyn_s.detach(),
        'z_sup': sy_from_quotient(z_sup_full[:, skip:]).detach(),
        'log_q': log_z_f.mean().detach(),
        'translik': trans.mean().detach(),
        'obj_appearances': app[:, skip:].detach() if app is not None else None,
    }
    if parts is not None:
        parts.update(lik_parts)
        parts.update(img_lik=img_lik, img_lik_sup=img_lik_sup, trans=trans, log_q=log_z_f)
    return average_elbo, prop, rewards


def supair_forward(c, P, x, noise, structs=None):
    """Stove.forward(..., pretrain=True) = bw_transform + Supair.forward, the SuPAIR-only ELBO
    (stove.py:863-897 pretrain branch, supair.py:504-551, sampling :165-192).

    noise: ONE standard-normal draw (n*T*O, 4) -- the rsample() of get_z_sup_sample.
    Returns (average_elbo, dict(z (n, T, O, 4), log_q, z_std))."""
    structs = structs or structures(c)
    if c.debug_bw:
        x = bw_transform(x)
    n, T = x.shape[0], x.shape[1]
    zp = encoder(c, P, x.flatten(end_dim=1)).flatten(end_dim=1)            # (nTO, 8)
    zp_mean, zp_std = constrain_zp(c, zp)
    eps = noise[0] if isinstance(noise, (list, tuple)) else noise
    z_tmp = zp_mean + zp_std * eps.to(zp_mean.dtype)                       # Normal(mean, std).rsample()
    log_q = normal_log_prob(z_tmp, zp_mean, zp_std).sum(-1)                # (nTO,)
    z_obj = sy_from_quotient(z_tmp)
    log_q = log_q.view(-1, c.num_obj).sum(-1)                              # (nT,)
    log_p = likelihood(c, P, structs, x, z_obj)
    log_p = log_p[0] if isinstance(log_p, tuple) else log_p
    elbo = log_p - log_q
    return elbo.mean(), dict(z=z_obj.view(n, T, c.num_obj, 4).detach(), log_q=log_q.mean().detach(),
                             z_std=zp_std.mean(0).detach())


def rollout(c, P, z_last, num=None, actions=None, appearance=None, noise=None,
            return_std=False):
    """stove.py:777-861.  `noise`: optional list of (n, O, cl//2) draws => sample=True."""
    cl = c.cl
    num = c.num_rollout if num is None else num
    z = [z_last]
    scale = z_last[..., :2]
    rewards, stds, log_qs = [], [], []
    if actions is not None:
        core_act, alen = actions.transpose(0, 1), actions.shape[1]
    else:
        core_act, alen = [None], 1
    for t in range(1, num + 1):
        tmp, reward = dynamics_forward(c, P, z[t - 1][..., 2:], 0, core_act[(t - 1) % alen],
                                       appearance)
        rewards.append(reward)
        zt, zstd = constrain_z_dyn(c, tmp[..., :cl // 2], tmp[..., cl // 2:])
        zt = torch.cat([z[t - 1][..., 2:4] + zt[..., :2], zt[..., 2:]], -1)
        stds.append(zstd)
        if noise is not None:
            mean = zt
            zt = mean + zstd * noise[t - 1]
            log_qs.append(normal_log_prob(zt, mean, zstd))
        z.append(torch.cat([scale, zt], -1))
    rewards = torch.stack(rewards, 1) if c.action_conditioned else torch.zeros(len(rewards))
    z_full = torch.stack(z[1:], 1)
    if noise is not None:
        return z_full, torch.stack(log_qs, 1), rewards
    if return_std:
        return z_full, torch.stack(stds, 1).detach(), rewards
    return z_full, rewards

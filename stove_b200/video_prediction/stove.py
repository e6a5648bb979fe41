"""STOVE model (drop-in for model/video_prediction/stove.py:11-897).

Same constructor (`Stove(config)`), `forward(x, step_counter, actions=None, pretrain=False)`
-> `(average_elbo, prop_dict, rewards)`, `rollout(...)`, `prop_dict` keys and `state_dict`
names as the reference.  The arithmetic of the hot path runs in the sm_100a kernels:
glimpse/masks + both SPNs inside `Supair.likelihood`, the interaction network inside
`Dynamics.forward`, and the whole time loop of `rollout` in one persistent kernel.
The sequence glue (matching, smoothing, Gaussian fusion, ELBO assembly) stays tensor code,
rewritten without the reference's per-step host synchronisation (stove.py:271-273) so a
training step can be captured in a CUDA graph.
"""
import math

import torch
import torch.nn as nn

from .dynamics import Dynamics
from .supair import Supair
from .. import ops
from ..utils.utils import bw_transform

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def _normal_log_prob(value, mean, std):
    return -((value - mean) ** 2) / (2 * std ** 2) - torch.log(std) - _HALF_LOG_2PI


class Stove(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.c = config
        self.step_counter = 0
        self.prop_dict = {}
        self.sup = Supair(config)
        self.dyn = Dynamics(config)
        self.reconstruct_from_z = self.sup.reconstruct_from_z
        kind = self.c.debug_match_objects
        if kind == '3_only':
            if self.c.num_obj != 3:
                raise ValueError('Matching Function not compatible w/ specified number of objects.')
            self.match_objects = self._3_only_match_objects
        elif kind == 'volatile':
            self.match_objects = self._volatile_match_objects
        elif kind == 'greedy':
            self.match_objects = self._greedy_match_objects
        else:
            raise ValueError('Specify valid self.c.debug_match_ojects.')
        if self.c.debug_no_latents or self.c.debug_no_velocity:
            # the fused dynamics loop (csrc/dynloop.cu) implements the default state fusion of stove.py:103-170 only
            raise NotImplementedError('stove_b200: debug_no_latents / debug_no_velocity (thesis ablations of '
                                      'Stove.full_state) are not implemented by the fused dynamics-loop kernels')

    def _sup_cfg(self, T):
        from .. import _native as N
        c = self.c
        key = ('sup_cfg', T)
        cache = self.__dict__.setdefault('_cfg_cache', {})
        if key not in cache:
            cache[key] = N.SupCfg(T, c.num_obj, ops.MATCH_KINDS[c.debug_match_objects],
                                  3 if (c.debug_core_appearance or c.debug_match_appearance) else 0,
                                  int(bool(c.debug_match_appearance)), int(bool(c.debug_fix_supair)),
                                  c.min_obj_scale, c.max_obj_scale, c.min_y_scale, c.max_y_scale,
                                  c.obj_pos_bound, c.scale_var, c.pos_var, 0.095)
        return cache[key]

    def _fuse_cfg(self):
        from .. import _native as N
        cache = self.__dict__.setdefault('_cfg_cache', {})
        if 'fuse' not in cache:
            std = [float(v) for v in self.dyn.transition_lik_std.flatten().tolist()]
            arr = (N.f32 * 32)(*(std + [1.0] * (32 - len(std))))
            cache['fuse'] = N.FuseCfg(self.c.pos_var, 0.04, self.c.debug_latent_q_std, arr)
        return cache['fuse']

    # -- noise: same shapes in the same order as the reference's rsample() calls ------------
    def _standard_normal(self, shape, like):
        return torch.empty(shape, device=like.device, dtype=like.dtype).normal_()

    def _standard_normal_n(self, count, shape, like):
        """`count` consecutive draws of `shape`, stacked on a new leading axis: one generator launch.
        A replacement `_standard_normal` that replays recorded draws (tests) sets `stacked = False`
        and is asked draw by draw, in the reference's order."""
        fn = self._standard_normal
        if getattr(fn, 'stacked', True):
            return fn((count,) + tuple(shape), like)
        draws = [fn(tuple(shape), like) for _ in range(count)]
        if like.is_cuda:
            # this runs on the packing side stream: tensors handed in from outside were allocated elsewhere and must
            # not return to the allocator before the copy below has read them
            here = torch.cuda.current_stream(like.device)
            for d in draws:
                if d.is_cuda:
                    d.record_stream(here)
        return torch.stack(draws, 0)

    # -- small sequence helpers ---------------------------------------------------------------
    def v_from_state(self, z_sup):
        """(n, T, o, 4) -> (n, T, o, 6): append finite-difference velocities, zeros at t=0."""
        full = torch.cat([z_sup[:, 1:], z_sup[:, 1:, :, 2:] - z_sup[:, :-1, :, 2:]], -1)
        return torch.cat([torch.zeros_like(full[:, :1]), full], 1)

    def v_std_from_pos(self, z_sup_std):
        v_std = torch.sqrt(z_sup_std[:, 1:, :, 2:] ** 2 + z_sup_std[:, :-1, :, 2:] ** 2)
        full = torch.cat([z_sup_std[:, 1:], v_std], -1)
        return torch.cat([torch.zeros_like(full[:, :1]), full], 1)

    def full_state(self, z_dyn, std_dyn, z_sup, std_sup):
        """Fuse the dynamics and SuPAIR Gaussians over (x, v), sample, score (stove.py:103-170)."""
        c = self.c
        mean_s, std_s = z_sup[..., :2], std_sup[..., :2]
        mean_l, std_l = z_dyn[..., 4:], std_dyn[..., 4:]
        m_sup, s_sup = z_sup[..., 2:6], std_sup[..., 2:6]
        m_dyn, s_dyn = z_dyn[..., :4], std_dyn[..., :4]
        v_sup, v_dyn = s_sup ** 2, s_dyn ** 2
        mean_xv = (v_sup * m_dyn + v_dyn * m_sup) / (v_dyn + v_sup)
        std_xv = s_dyn * s_sup / torch.sqrt(v_dyn + v_sup)
        if c.debug_no_latents:
            mean = torch.cat([mean_s, mean_xv], -1)
            std = torch.cat([std_s, std_xv], -1)
            z_s = mean + std * self._standard_normal(mean.shape, mean)
            log_q = _normal_log_prob(z_s, mean, std)
            return torch.cat([z_s, torch.zeros_like(mean_l)], -1), log_q, mean, std
        if c.debug_no_velocity:
            mean = torch.cat([mean_s, mean_xv[..., :2], torch.zeros_like(mean_xv[..., 2:]), mean_l], -1)
            std = torch.cat([std_s, std_xv[..., :2], torch.ones_like(std_xv[..., 2:]), std_l], -1)
        else:
            mean = torch.cat([mean_s, mean_xv, mean_l], -1)
            std = torch.cat([std_s, std_xv, std_l], -1)
        z_s = mean + std * self._standard_normal(mean.shape, mean)
        return z_s, _normal_log_prob(z_s, mean, std), mean, std

    def transition_lik(self, means, results):
        return _normal_log_prob(results, means, self.dyn.transition_lik_std.to(results.dtype))

    # -- object matching ------------------------------------------------------------------------
    def _match_inputs(self, z_sup, z_sup_std, obj_appearances):
        z = (z_sup + 1) / 2
        m_idx = slice(2, 4)                 # a slice, not an index list: no H2D copy under graph capture
        if obj_appearances is not None:
            z = torch.cat([z, obj_appearances], -1)
            if self.c.debug_match_appearance:
                m_idx = slice(2, 7)
        if z_sup_std is not None:
            z = torch.cat([z, z_sup_std], -1)
        return z, m_idx

    @staticmethod
    def _match_outputs(z_matched, z_sup_std, obj_appearances):
        z_sup_matched = 2 * z_matched[..., :4] - 1
        app = z_matched[..., 4:7] if obj_appearances is not None else None
        if z_sup_std is None and obj_appearances is not None:
            return z_sup_matched, app
        if z_sup_std is not None and obj_appearances is None:
            return z_sup_matched, z_matched[..., 4:8], None
        if z_sup_std is not None and obj_appearances is not None:
            return z_sup_matched, z_matched[..., 7:11], app
        return z_sup_matched

    @staticmethod
    def _pair_errors(prev, curr):
        """err[b, a, j] = |prev_a - curr_j|^2 (detached)."""
        return ((prev.detach().unsqueeze(2) - curr.detach().unsqueeze(1)) ** 2).sum(-1)

    def _3_only_match_objects(self, z_sup, z_sup_std=None, obj_appearances=None):
        """Nearest-neighbour matching with greedy repair of non-permutations (stove.py:200-329),
        without the reference's `if num_faults > 0` host sync: the repair is evaluated for every
        row and selected with a mask."""
        z, m_idx = self._match_inputs(z_sup, z_sup_std, obj_appearances)
        O = self.c.num_obj
        matched = [z[:, 0]]
        for t in range(1, z.shape[1]):
            err = self._pair_errors(matched[t - 1][..., m_idx], z[:, t][..., m_idx])
            idx = err.argmin(-1)
            valid = (idx[:, 0] != idx[:, 1]) & (idx[:, 1] != idx[:, 2]) & (idx[:, 0] != idx[:, 2])
            e = err
            fixed = []
            for o in range(O):
                col = e.argmin(-1)[:, o]
                fixed.append(col)
                e = e.masked_fill(torch.nn.functional.one_hot(col, O).bool().unsqueeze(1), 1e12)
            idx = torch.where(valid.unsqueeze(1), idx, torch.stack(fixed, 1))
            matched.append(torch.gather(z[:, t], 1, idx.unsqueeze(-1).expand(-1, -1, z.shape[-1])))
        return self._match_outputs(torch.stack(matched, 1), z_sup_std, obj_appearances)

    def _volatile_match_objects(self, z_sup, z_sup_std=None, obj_appearances=None):
        """Per-object nearest neighbour, no permutation check (stove.py:331-430)."""
        z, m_idx = self._match_inputs(z_sup, z_sup_std, obj_appearances)
        matched = [z[:, 0]]
        for t in range(1, z.shape[1]):
            err = self._pair_errors(matched[t - 1][..., m_idx], z[:, t][..., m_idx])   # [b, prev, cur]
            col = err.argmin(-1)
            matched.append(torch.gather(z[:, t], 1, col.unsqueeze(-1).expand(-1, -1, z.shape[-1])))
        return self._match_outputs(torch.stack(matched, 1), z_sup_std, obj_appearances)

    def _greedy_match_objects(self, z_sup, z_sup_std=None, obj_appearances=None):
        """Global greedy bipartite matching (stove.py:432-514)."""
        z, m_idx = self._match_inputs(z_sup, z_sup_std, obj_appearances)
        O = self.c.num_obj
        matched = [z[:, 0]]
        for t in range(1, z.shape[1]):
            err = self._pair_errors(matched[t - 1][..., m_idx], z[:, t][..., m_idx])
            n = err.shape[0]
            assign = torch.zeros(n, O, dtype=torch.long, device=z.device)
            for _ in range(O):
                flat = err.view(n, -1).argmin(1)
                ix, iy = flat // O, flat % O
                assign = torch.where(torch.nn.functional.one_hot(ix, O).bool(), iy.unsqueeze(1), assign)
                big = err.max() + 1
                err = err.masked_fill(torch.nn.functional.one_hot(ix, O).bool().unsqueeze(2), 0) \
                    + torch.nn.functional.one_hot(ix, O).unsqueeze(2).to(err.dtype) * big
                big = err.max() + 1
                err = err.masked_fill(torch.nn.functional.one_hot(iy, O).bool().unsqueeze(1), 0) \
                    + torch.nn.functional.one_hot(iy, O).unsqueeze(1).to(err.dtype) * big
            matched.append(torch.gather(z[:, t], 1, assign.unsqueeze(-1).expand(-1, -1, z.shape[-1])))
        return self._match_outputs(torch.stack(matched, 1), z_sup_std, obj_appearances)

    def fix_supair(self, z, z_std=None):
        """Replace states that jump away from both temporal neighbours by their average
        (stove.py:516-571)."""
        zz = torch.cat([z, z_std], -1) if z_std is not None else z
        d = (zz[:, 1:, :, :2] - zz[:, :-1, :, :2]).abs().detach()
        zero = torch.zeros_like(d[:, :1])
        flag = (torch.cat([zero, d], 1) > 0.095) & (torch.cat([d, zero], 1) > 0.095)
        pad = torch.zeros_like(zz[:, :1])
        smooth = torch.cat([pad, (zz[:, :-2] + zz[:, 2:]) / 2, pad], 1)
        zz = torch.where(torch.cat(zz.shape[-1] // 2 * [flag], -1), smooth, zz)
        if z_std is not None:
            return torch.chunk(zz, 2, dim=-1)
        return zz

    def object_embedding(self, z, x_color):
        """Mean colour of each object's glimpse (stove.py:573-597)."""
        z_patch = self.sup.sy_from_quotient(z[..., :4].detach())
        patches = self.sup.patches_from_z(x_color.flatten(end_dim=1), z_patch.flatten(end_dim=2))
        return patches.mean((-1, -2)).view(*z.shape[:-1], 3)

    # -- sequence ELBO ----------------------------------------------------------------------------
    def stove_forward(self, x, actions=None, x_color=None, x_planes=None):
        c = self.c
        n, T = x.shape[0], x.shape[1]
        skip, cl, O = c.skip, c.cl, c.num_obj
        # Everything that does not depend on the encoder output runs on a side stream next to the encoder and is
        # joined right before the dynamics loop: parameter packing (~20 short launches, and its backward), the
        # noise of the step and the contiguous copy of the scored frames x[:, 1:].  The encoder is ISSUED first:
        # in the captured step, ready graph nodes are launched in creation order, and the short side-stream
        # launches would otherwise delay the first kernel of the chain by ~10 us.
        cur = torch.cuda.current_stream(x.device)
        pack_stream = self.sup._side_stream(x.device, 'pack')
        forked = torch.cuda.Event()
        forked.record(cur)

        # encoder -> (constrain, match, smooth, velocities) in one kernel (csrc/glue.cu)
        zp = self.sup.encoder(x.flatten(end_dim=1), planes=x_planes).view(n, T, O, 8)

        # the dynamics weights on a stream of their own: their backward (and the weight-gradient kernels that
        # DynamicsLoop.backward joins into this stream) must not sit in front of the backward of the SPN packing
        dyn_stream = self.sup._side_stream(x.device, 'pack_dyn') if ops.dyn_stream_enabled() else pack_stream
        pack_stream.wait_event(forked)
        if dyn_stream is not pack_stream:
            dyn_stream.wait_event(forked)
        with torch.cuda.stream(dyn_stream):
            packed_dyn = self.dyn.pack_weights(0, actions is not None, c.debug_core_appearance)
        with torch.cuda.stream(pack_stream):
            packed_spn = self.sup.pack()
            # initial latents ~ N(0, 0.01^2) (stove.py:672-680).  The reference draws a second sample of the same
            # shape for the initial dynamics std (logging only); both come from one generator launch
            prior_shape = (n, O, cl // 2 - 4, 1)
            pri = self._standard_normal_n(2, prior_shape, x)
            lat0 = (0.01 * pri[0]).squeeze(-1)
            eps = self._standard_normal_n(T - skip, (n, O, cl // 2 + 2), x)
            x_scored = x[:, 1:].flatten(end_dim=1).contiguous()

        _app = None
        if c.debug_core_appearance or c.debug_match_appearance:
            with torch.no_grad():
                z_raw, _ = self.sup.constrain_zp(zp.flatten(end_dim=2))
            _app = self.object_embedding(z_raw.view(n, T, O, 4), x_color)
        z_sup, z_sup_full, z_sup_std_full, obj_appearances = ops.SupPrepare.apply(zp, _app, self._sup_cfg(T))
        if _app is None:
            obj_appearances = None

        # dynamics loop: the whole loop is one persistent kernel (csrc/dynloop.cu), chained on the device
        cur.wait_stream(pack_stream)
        if dyn_stream is not pack_stream:
            cur.wait_stream(dyn_stream)
        cfg_dyn, w_dyn = packed_dyn
        for t in (w_dyn, lat0, eps, x_scored) + tuple(v for pk in packed_spn if pk is not None
                                                        for v in vars(pk).values() if isinstance(v, torch.Tensor)):
            t.record_stream(cur)
        z_s, z_dyn_s, z_dyn_std_s, z_std_s, log_z_n, trans_n, rewards = ops.DynamicsLoop.apply(
            z_sup_full, z_sup_std_full, lat0, eps, actions,
            obj_appearances if c.debug_core_appearance else None, w_dyn, cfg_dyn, self._fuse_cfg(), skip,
            dyn_stream)
        if not c.action_conditioned:
            rewards = torch.zeros(T - skip)

        # p(x_t | z_t) for t >= skip and p(x_t | z_sup_t) for 1 <= t < skip (stove.py:731-736) share one
        # pass over the frames x[:, 1:]: a single glimpse/mask launch and one launch family per SPN;
        # the per-frame terms are reduced to the ELBO by one kernel (csrc/glue.cu)
        fused = self.sup.sequence_elbo(x_scored, z_sup, z_s, log_z_n, trans_n, skip, packed_spn)
        if fused is not None:
            # one launch each way: states read from z_sup / z_s, ELBO assembled by the kernel (ops.SceneElbo)
            average_elbo, stats, bg, patch_raw, overlap, extra = fused
            z_all = None
        else:
            z_all = ops.ZAll.apply(z_sup, z_s, skip)                                # (n, T-1, O, 4) [sx, sy, x, y]
            bg, patch_raw, overlap, extra = self.sup.likelihood_raw(x_scored,
                                                                    z_all.flatten(end_dim=1), packed=packed_spn)
            average_elbo, stats = ops.ElboAssemble.apply(bg, patch_raw, z_all, overlap, log_z_n, trans_n, skip,
                                                         float(c.overlap_beta))
        logging = (self.step_counter % c.print_every == 0) or (self.step_counter % c.plot_every == 0)
        if logging and c.debug:
            p = self.prop_dict
            p['bg'], p['patch'], p['overlap'] = stats[1], stats[2], stats[3]
            if c.debug_extend_plots:
                n_sup = skip - 1
                with torch.no_grad():
                    if z_all is None:
                        z_all = ops.ZAll.apply(z_sup.detach(), z_s.detach(), skip)
                    za = z_all.flatten(end_dim=1)
                    pl = (patch_raw.view(-1, O) * za[..., 0] * za[..., 1]).sum(1)
                p['overlap_ratios'] = extra['overlap_ratios'].detach()
                p['patches'] = extra['patches'].detach()
                p['marginalise_flat'] = extra['marginalise_flat'].detach()
                p['patches_loglik'] = pl.view(n, T - 1)[:, n_sup:].flatten()
                p['marginalise_bg'] = extra['marginalise_bg'].detach()
                p['bg_loglik'] = bg.detach().view(n, T - 1)[:, n_sup:].flatten()

        if (self.step_counter % c.print_every == 0) or (self.step_counter % c.plot_every == 0):
            p = self.prop_dict
            p['z'] = self.sup.sy_from_quotient(z_s).detach()
            p['z_dyn'] = z_dyn_s.detach()
            p['z_sup'] = self.sup.sy_from_quotient(z_sup_full[:, skip:]).detach()
            p['z_std'] = z_std_s.mean((0, 1, 2)).detach()
            p['z_dyn_std'] = torch.cat([z_s.new_full((2,), float('nan')),
                                        z_dyn_std_s[..., :4].mean((0, 1, 2)).detach()])
            p['z_sup_std'] = z_sup_std_full[:, skip:].mean((0, 1, 2)).detach()
            p['log_q'] = stats[4]
            p['translik'] = stats[5]
            p['obj_appearances'] = obj_appearances[:, skip:].detach() if obj_appearances is not None else None
            if c.debug and c.debug_extend_plots:
                p['z_dyn_std_full'] = z_dyn_std_s.detach()
        return average_elbo, self.prop_dict, rewards

    # -- rollout ------------------------------------------------------------------------------------
    @torch.no_grad()
    def rollout(self, z_last, num=None, sample=False, return_std=False, actions=None, appearance=None):
        """Roll the dynamics model out for `num` steps in ONE persistent kernel
        (stove.py:777-861).  z_last (n, o, cl//2 + 2) holds [sx, sy | x, y | v | latent].
        Returns (z_full (n, num, o, cl//2+2), rewards) -- or the 3-tuples of the reference for
        `sample` / `return_std`.  Outputs are detached (inference path)."""
        c = self.c
        num = c.num_rollout if num is None else num
        cfg, weights = self.dyn.pack_weights(0, actions is not None, appearance is not None)
        noise = None
        if sample:
            noise = self._standard_normal((z_last.shape[0], num, c.num_obj, c.cl // 2), z_last)
        z_full, std, logq, rewards = ops.gnn_rollout(
            cfg, z_last, num, weights, actions=actions, app=appearance, noise=noise, pos_var=c.pos_var,
            vel_std=0.04, latent_std=c.debug_latent_q_std, want_std=return_std)
        if not c.action_conditioned:
            rewards = torch.zeros(num)
        if sample:
            return z_full, logq, rewards
        if return_std:
            return z_full, std, rewards
        return z_full, rewards

    # -- entry point ------------------------------------------------------------------------------------
    def forward(self, x, step_counter, actions=None, pretrain=False):
        """x (n, T, 3, w, h) in [0, 1]; returns (average_elbo, prop_dict, rewards) [stove.py:863-897]."""
        self.step_counter = step_counter
        self.sup.step_counter = step_counter
        self.dyn.step_counter = step_counter
        x_color, x_planes = x, None
        if isinstance(x, ops.IndirectFrames) and (pretrain or not self.c.debug_bw or self.c.debug_core_appearance
                                                  or self.c.debug_match_appearance):
            raise ValueError('IndirectFrames are only consumed by bw_transform (debug_bw, no appearances, no pretrain)')
        if not pretrain:
            self.sup.encoder.prepare()          # frame-independent encoder work, off the chain (side stream)
        if self.c.debug_bw:
            # one pass: channel sum + clamp (uint8 frames are scaled by 1/255 here) and, for the sequence model,
            # the TF32 operand planes of the frames for the recognition LSTM's input GEMM
            if pretrain or not x.is_cuda:
                x = bw_transform(x)
            else:
                x, x_planes = ops.bw_transform(x, want_planes=True)
        elif x.dtype == torch.uint8:
            x = x.float() / 255
        if pretrain:
            elbo, prop_dict = self.sup(x)
            return elbo, prop_dict, 0
        if self.c.debug_core_appearance or self.c.debug_match_appearance:
            if x_color.dtype == torch.uint8:
                x_color = x_color.float() / 255
            return self.stove_forward(x, actions=actions, x_color=x_color, x_planes=x_planes)
        return self.stove_forward(x, actions=actions, x_planes=x_planes)

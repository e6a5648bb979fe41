"""SuPAIR scene model (drop-in for model/video_prediction/supair.py:14-551).

`likelihood` = background-SPN(frame | background mask) + sum_obj object-SPN(glimpse | overlap
mask) * sx * sy + Exponential(beta) prior on the overlap ratios (supair.py:44-110).  The
glimpses and the sequential compositing masks come from one kernel per call
(csrc/scene.cu), the two SPNs from the fused kernels in csrc/spn_obj.cu / spn_bg.cu.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import encoder
from .. import ops
from ..spn import probabilistic_models as prob


class _SimpleGauss:
    """Fixed single-Gaussian stand-ins used by the reference's debug flags
    (probabilistic_models.py:42-91)."""

    def __init__(self, mean, var):
        self.mean, self.var = mean, var

    def forward(self, flat, marg_flat):
        ll = -((flat - self.mean) ** 2) / (2 * self.var ** 2) - math.log(self.var) - 0.5 * math.log(2 * math.pi)
        return (ll * (1 - marg_flat)).sum(1).unsqueeze(-1)


class Supair(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.c = config
        self.step_counter = 0
        self.prop_dict = {}
        self.encoder = encoder.RnnStates(self.c)
        if self.c.debug_obj_spn:
            self.obj_spn = _SimpleGauss(0.8, self.c.debug_simple_obj_var)
        else:
            self.obj_spn = prob._get_obj_spn(self.c, seed=self.c.random_seed)
        if self.c.debug_bg_model:
            self.bg_spn = _SimpleGauss(0.0, self.c.debug_simple_bg_var)
        else:
            self.bg_spn = prob._get_bg_spn(self.c, seed=self.c.random_seed)

    # -- packing -------------------------------------------------------------------------
    def pack(self):
        """Kernel-layout parameters of both SPNs; pack once, score many frame sets."""
        return (self.obj_spn.pack() if isinstance(self.obj_spn, nn.Module) else None,
                self.bg_spn.pack() if isinstance(self.bg_spn, nn.Module) else None)

    def _align(self):
        return bool(getattr(self.c, 'align_corners', False))

    def _const(self, key, values, like):
        """Small constant tensors, cached per device/dtype (no H2D copy inside a captured step)."""
        cache = self.__dict__.setdefault('_const_cache', {})
        k = (key, like.device, like.dtype)
        if k not in cache:
            cache[k] = torch.tensor(values, device=like.device, dtype=like.dtype)
        return cache[k]

    def _side_stream(self, device, name='bg'):
        """A second CUDA stream (per device) for work that is independent of the main chain: the
        background SPN runs there while the object SPN runs on the current stream -- both are
        short, latency-bound launches that leave most SMs idle on their own.  autograd replays
        each op's backward on the stream of its forward, so the two backward chains overlap too;
        a CUDA-graph capture records the fork/join as parallel branches."""
        if not ops.fork_enabled():                   # serial execution (per-kernel timing passes)
            return torch.cuda.current_stream(device)
        cache = self.__dict__.setdefault('_streams', {})
        # 'bg' carries kernels the main chain waits for: same priority as the caller's stream; the others
        # ('pack') are off the chain: default (= lowest) priority
        prio = torch.cuda.current_stream(device).priority if name == 'bg' else 0
        key = (device.type, device.index, name, prio)
        if key not in cache:
            cache[key] = torch.cuda.Stream(device=device, priority=prio)
        return cache[key]

    # -- likelihood ----------------------------------------------------------------------
    def likelihood_raw(self, x_img, z_img, packed=None):
        """x_img (F, c, w, h), z_img (F, O, 4) [sx, sy, x, y] -> background log-likelihood (F,), RAW
        object-SPN log-likelihoods (F*O,) (not yet weighted by sx * sy), overlap ratios (F, O) and the
        intermediate tensors: one glimpse/mask launch, one launch family per SPN."""
        c = self.c
        pk_obj, pk_bg = packed if packed is not None else self.pack()
        if pk_obj is not None and pk_bg is not None and ops.scene_ll_supported(
                x_img, z_img, c.patch_width, c.patch_height, pk_obj.tables, pk_bg.tables):
            # one launch for the whole likelihood (csrc/scene_ll.cu): glimpses and masks never leave the SM
            cur = torch.cuda.current_stream(x_img.device)
            streams = [None if p.stream is None or p.stream == cur else p.stream for p in (pk_obj, pk_bg)]
            bg_loglik, patches_loglik, overlap, patches, marg_patch, marg_bg = ops.SceneLL.apply(
                x_img, z_img, pk_obj.leaf, pk_obj.wlog, pk_obj.wlin, pk_obj.rlog, pk_obj.rlin,
                pk_bg.leaf, pk_bg.rlog, pk_bg.rlin, pk_obj.tables, pk_bg.tables,
                c.patch_width, c.patch_height, self._align(), streams[0], streams[1], pk_bg.leaf_il_f, pk_bg.leaf_il_b)
            extra = dict(overlap_ratios=overlap, patches=patches, marginalise_flat=marg_patch.flatten(start_dim=1),
                         marginalise_bg=marg_bg)
            return bg_loglik, patches_loglik, overlap, extra
        patches, marg_patch, marg_bg, overlap = ops.Scene.apply(
            x_img, z_img, c.patch_width, c.patch_height, self._align())
        img_flat, marg_flat = x_img.flatten(start_dim=1), marg_bg.flatten(start_dim=1)
        cur = torch.cuda.current_stream(x_img.device)
        side = self._side_stream(x_img.device)
        forked = torch.cuda.Event()
        forked.record(cur)
        # the object SPN is the longer of the two and is issued first: in the captured step the launch order of
        # ready graph nodes follows their creation order, and the background kernels fill every SM
        patches_flat, marginalise_flat = patches.flatten(start_dim=1), marg_patch.flatten(start_dim=1)
        if pk_obj is not None:
            patches_loglik = self.obj_spn.forward_packed(pk_obj, patches_flat, marginalise_flat)[:, 0]
        else:
            patches_loglik = self.obj_spn.forward(patches_flat, marginalise_flat)[:, 0]
        side.wait_event(forked)
        with torch.cuda.stream(side):
            if pk_bg is not None:
                bg_loglik = self.bg_spn.forward_packed(pk_bg, img_flat, marg_flat)[:, 0]
            else:
                bg_loglik = self.bg_spn.forward(img_flat, marg_flat)[:, 0]
        cur.wait_stream(side)
        bg_loglik.record_stream(cur)
        extra = dict(overlap_ratios=overlap, patches=patches, marginalise_flat=marginalise_flat,
                     marginalise_bg=marg_bg)
        return bg_loglik, patches_loglik, overlap, extra

    def sequence_elbo(self, x_img, z_sup, z_s, logq, trans, skip, packed):
        """The likelihood half of the sequence ELBO (stove.py:731-748) in one launch each way (ops.SceneElbo), or
        None when the configuration has no fused kernel.  Returns (elbo, stats, bg, raw object ll, overlap, extra)."""
        c = self.c
        pk_obj, pk_bg = packed
        if not ops.scene_seq_enabled() or pk_obj is None or pk_bg is None or pk_bg.leaf_il_f is None or not ops.scene_ll_supported(
                x_img, z_s[:, 0], c.patch_width, c.patch_height, pk_obj.tables, pk_bg.tables):
            return None
        cur = torch.cuda.current_stream(x_img.device)
        streams = [None if p.stream is None or p.stream == cur else p.stream for p in (pk_obj, pk_bg)]
        elbo, stats, bg, obj, overlap, patches, marg_patch, marg_bg = ops.SceneElbo.apply(
            x_img, z_sup, z_s, logq, trans, skip, float(c.overlap_beta), pk_obj.leaf, pk_obj.wlog, pk_obj.wlin,
            pk_obj.rlog, pk_obj.rlin, pk_bg.leaf, pk_bg.rlog, pk_bg.rlin, pk_obj.tables, pk_bg.tables,
            c.patch_width, c.patch_height, self._align(), streams[0], streams[1], pk_bg.leaf_il_f, pk_bg.leaf_il_b)
        extra = dict(overlap_ratios=overlap, patches=patches, marginalise_flat=marg_patch.flatten(start_dim=1),
                     marginalise_bg=marg_bg)
        return elbo, stats, bg, obj, overlap, extra

    def likelihood_parts(self, x_img, z_img, packed=None):
        """Per-frame (bg, patch, overlap) log-likelihood terms of supair.py:84-110 plus the
        intermediate tensors."""
        c = self.c
        bg_loglik, patches_loglik, overlap, extra = self.likelihood_raw(x_img, z_img, packed)
        z_flat = z_img.reshape(-1, 4)
        patches_loglik = (patches_loglik * z_flat[:, 0] * z_flat[:, 1]).view(-1, c.num_obj).sum(1)
        # log Exponential(beta)(overlap) = log beta - beta * overlap
        overlap_log_liks = (math.log(c.overlap_beta) - c.overlap_beta * overlap).sum(1)
        return bg_loglik, patches_loglik, overlap_log_liks, extra

    def _log_parts(self, bg_loglik, patches_loglik, overlap_log_liks, extra):
        c = self.c
        if (self.step_counter % c.print_every == 0) or (self.step_counter % c.plot_every == 0):
            if c.debug:
                self.prop_dict['bg'] = bg_loglik.mean().detach()
                self.prop_dict['patch'] = patches_loglik.mean().detach()
                self.prop_dict['overlap'] = overlap_log_liks.mean().detach()
            if c.debug and c.debug_extend_plots:
                self.prop_dict['overlap_ratios'] = extra['overlap_ratios'].detach()
                self.prop_dict['patches'] = extra['patches'].detach()
                self.prop_dict['marginalise_flat'] = extra['marginalise_flat'].detach()
                self.prop_dict['patches_loglik'] = patches_loglik.detach()
                self.prop_dict['marginalise_bg'] = extra['marginalise_bg'].detach()
                self.prop_dict['bg_loglik'] = bg_loglik.detach()

    def likelihood(self, x, z_obj, packed=None):
        """x (n, T, c, w, h), z_obj (n*T*O, 4) [sx, sy, x, y] -> (log p(x|z) (n*T,), prop_dict)
        [supair.py:44-110]."""
        x_img = x.flatten(end_dim=1)
        bg, patch, ov, extra = self.likelihood_parts(x_img, z_obj.view(-1, self.c.num_obj, 4), packed)
        self._log_parts(bg, patch, ov, extra)
        return bg + patch + ov, self.prop_dict

    # -- z handling ----------------------------------------------------------------------
    def constrain_zp(self, zp):
        """(nTo, 8) raw encoder output -> means (sx, sy/sx, x, y) and stds, supair.py:112-149."""
        c = self.c
        sig = torch.sigmoid(zp)
        hi = self._const('hi', [c.max_obj_scale - c.min_obj_scale, c.max_y_scale - c.min_y_scale,
                                2 * c.obj_pos_bound, 2 * c.obj_pos_bound], zp)
        lo = self._const('lo', [c.min_obj_scale, c.min_y_scale, -c.obj_pos_bound, -c.obj_pos_bound], zp)
        std_scale = self._const('std', [c.scale_var, c.scale_var, c.pos_var, c.pos_var], zp)
        return sig[:, :4] * hi + lo, sig[:, 4:8] * std_scale

    @staticmethod
    def sy_from_quotient(z):
        return torch.cat([z[..., 0:1], z[..., 0:1] * z[..., 1:2], z[..., 2:]], -1)

    @staticmethod
    def quotient_from_sy(z):
        return torch.cat([z[..., 0:1], z[..., 1:2] / z[..., 0:1], z[..., 2:]], -1)

    def _standard_normal(self, shape, like):
        """Noise source of get_z_sup_sample (replaced by tests that replay recorded draws)."""
        return torch.empty(shape, device=like.device, dtype=like.dtype).normal_()

    def get_z_sup_sample(self, zp_mean, zp_std):
        """supair.py:165-192: z ~ N(zp_mean, zp_std) (reparametrised), log q(z) summed over the 4 components."""
        eps = self._standard_normal(tuple(zp_mean.shape), zp_mean)
        z = zp_mean + zp_std * eps
        log_q = (-0.5 * eps ** 2 - torch.log(zp_std) - 0.5 * math.log(2 * math.pi)).sum(-1)
        return self.sy_from_quotient(z), log_q

    @staticmethod
    def expand_z(z):
        """[sx, sy, x, y] -> [[sx, 0, x], [0, sy, y]]."""
        zero = torch.zeros_like(z[:, 0])
        return torch.stack([z[:, 0], zero, z[:, 2], zero, z[:, 1], z[:, 3]], 1).view(-1, 2, 3)

    @staticmethod
    def invert_z(z):
        return torch.stack([1. / z[:, 0], 1. / z[:, 1], -z[:, 2] / z[:, 0], -z[:, 3] / z[:, 1]], 1)

    def patches_from_z(self, x_img, z_obj):
        """x_img (nT, c, w, h), z_obj (nTo, 4) -> glimpses (nTo, c, pw, ph)  [supair.py:241-276]."""
        z_img = z_obj.reshape(x_img.shape[0], -1, 4)
        return ops.Scene.apply(x_img, z_img, self.c.patch_width, self.c.patch_height, self._align())[0]

    def masks_from_z(self, z_img):
        """z_img (F, O, 4) -> (marg_patch (F*O, c, pw, ph), marg_bg (F, c, w, h), overlap (F, O))."""
        c = self.c
        blank = z_img.new_zeros(z_img.shape[0], c.channels, c.width, c.height)
        _, mp, mb, ov = ops.Scene.apply(blank, z_img, c.patch_width, c.patch_height, self._align())
        return mp, mb, ov

    # -- visualisation helpers (not on the hot path; plain torch) --------------------------
    def spn_max_activation(self, spn=None):
        spn = spn if spn is not None else self.obj_spn
        idx = {vec: np.argmax(p.detach().cpu().numpy(), 0) for vec, p in spn.get_sum_params().items()}
        img = np.clip(spn.reconstruct(idx, 0, sample=False), 0., 1.)
        ref = spn.output_vector.params
        return torch.as_tensor(img, device=ref.device, dtype=ref.dtype)

    def spn_mpe(self, z, x, spn=None):
        spn = self.bg_spn if spn == 'bg' else (spn if spn is not None else self.obj_spn)
        if x.shape[0] != z.shape[0]:
            raise ValueError('x and z need to have same batch_dim.')
        patches = self.patches_from_z(x, z.flatten(end_dim=1))
        _, child = spn.compute_activations(patches.flatten(start_dim=1), get_sum_child_acts=True)
        recons = []
        for j in range(x.shape[0] * self.c.num_obj):
            idx = {vec: np.argmax(p[j].detach().cpu().numpy(), 0) for vec, p in child.items()}
            recons.append(torch.as_tensor(spn.reconstruct(idx, 0, False), device=x.device, dtype=x.dtype))
        return torch.stack(recons, 0).clamp(0., 1.).view(x.shape[0], self.c.num_obj, -1)

    def reconstruct_from_z(self, z, x=None, max_activation=True, single_image=True):
        """Render states z (n, T, o, >=4) with the SPNs' most probable appearance (supair.py:425-501).  On the
        GPU the paste loop is one kernel (ops.render); CPU tensors take the reference's grid_sample loop."""
        import torch.nn.functional as F
        c = self.c
        z = z[..., :4]
        ph, pw, h, w = c.patch_height, c.patch_width, c.height, c.width
        n_img = z.shape[0] * z.shape[1]
        canvas = self.spn_max_activation(self.bg_spn).view(1, 1, w, h).repeat(n_img, 1, 1, 1)
        if max_activation:
            obj = self.spn_max_activation(self.obj_spn).view(1, 1, c.channels, pw, ph)
            obj = obj.repeat(n_img, c.num_obj, 1, 1, 1)
        else:
            if x is None:
                raise ValueError('Need x for reconstructions.')
            z_in, x_in = (z[:, 0], x) if single_image else (z.flatten(end_dim=1), x.flatten(end_dim=1))
            obj = self.spn_mpe(z_in, x_in, self.obj_spn).view(z_in.shape[0], c.num_obj, c.channels, pw, ph)
            if single_image:
                obj = obj.unsqueeze(1).repeat(1, z.shape[1], 1, 1, 1, 1).flatten(end_dim=1)
        z_img = z.flatten(end_dim=1)
        if z_img.is_cuda:
            # one kernel for the whole paste loop (csrc/scene.cu: render_kernel)
            out = ops.render(canvas, obj, z_img, w, h, self._align())
            return out.view(*z.shape[:2], c.channels, w, h)
        ac = self._align()
        for o in range(c.num_obj):
            theta = self.expand_z(self.invert_z(z_img[:, o]))
            grid = F.affine_grid(theta, torch.Size((n_img, 1, w, h)), align_corners=ac)
            canvas = canvas + F.grid_sample(obj[:, o], grid, align_corners=ac)
        return canvas.view(*z.shape[:2], c.channels, w, h).clamp(0, 1)

    # -- SuPAIR-only ELBO (pretraining) -----------------------------------------------------
    def forward(self, x):
        """x (n, T, c, w, h) -> (mean ELBO of SuPAIR alone, prop_dict)   [supair.py:504-551]."""
        c = self.c
        zp = self.encoder(x.flatten(end_dim=1)).flatten(end_dim=1)
        zp_mean, zp_std = self.constrain_zp(zp)
        z_obj, log_q = self.get_z_sup_sample(zp_mean, zp_std)
        log_q = log_q.view(-1, c.num_obj).sum(-1)
        log_p, _ = self.likelihood(x, z_obj)
        elbo = log_p - log_q
        average_elbo = elbo.mean()
        if (self.step_counter % c.print_every == 0) or (self.step_counter % c.plot_every == 0):
            self.prop_dict['z'] = z_obj.view(*x.shape[0:2], c.num_obj, 4).detach()
            if c.debug:
                self.prop_dict['log_q'] = log_q.mean().detach()
                self.prop_dict['z_std'] = zp_std.mean(0).detach()
            if c.debug and c.debug_extend_plots:
                self.prop_dict['elbo'] = elbo.detach()
                self.prop_dict['log_q_xz'] = log_q.detach()
        return average_elbo, self.prop_dict

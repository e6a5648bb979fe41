"""Graph-network dynamics (drop-in for model/video_prediction/dynamics.py:8-265).

Same constructor, parameter names and call signature; `forward` runs the fused sm_100a
interaction-network kernel (stove_b200/csrc/gnn.cu) instead of ~370 ATen ops.
"""
import torch
import torch.nn as nn

from .. import _native as N
from .. import ops


class Dynamics(nn.Module):
    def __init__(self, config, enc_input_size=None):
        super().__init__()
        self.c = config
        self.step_counter = 0
        self.prop_dict = {}
        cl = self.c.cl
        # enc_input_size: width of the state rows `forward` receives (dynamics.py:24-26); the supervised
        # ablation passes 16 with lim_enc = 4 (supairvised/dynamics.py:24-25, 75-77).  Any width up to cl runs
        # through the single-step kernels; the fused loop / rollout kernels need the STOVE layout (cl // 2).
        self.enc_input_size = cl // 2 if enc_input_size is None else int(enc_input_size)
        if not 0 < self.enc_input_size <= cl:
            raise ValueError('enc_input_size must be in (0, cl]')
        enc_in = self.enc_input_size
        if self.c.action_conditioned:
            self.n_action_enc = 4
            self.action_embedding_layer = nn.Linear(self.c.action_space, self.c.num_obj * self.n_action_enc)
            enc_in += self.n_action_enc
            self.reward_head0 = nn.Sequential(nn.Linear(cl, cl), nn.ReLU(), nn.Linear(cl, cl))
            self.reward_head1 = nn.Sequential(nn.Linear(cl, cl // 2), nn.ReLU(), nn.Linear(cl // 2, cl // 4),
                                              nn.ReLU(), nn.Linear(cl // 4, 1))
        if self.c.debug_core_appearance:
            enc_in += self.c.debug_appearance_dim
        self.state_enc = nn.Linear(enc_in, cl)

        def stack(sizes):
            return nn.ModuleList(nn.ModuleList(nn.Linear(i, o) for i, o in sizes) for _ in range(3))

        self.self_cores = stack([(cl, cl), (cl, cl)])
        self.rel_cores = stack([(1 + 2 * cl, 2 * cl), (2 * cl, cl), (cl, cl)])
        self.att_net = stack([(1 + 2 * cl, 2 * cl), (2 * cl, cl), (cl, 1)])
        self.affector = stack([(cl, cl), (cl, cl), (cl, cl)])
        self.out = stack([(cl + cl, cl), (cl, cl)])
        if self.c.debug_xavier:
            self.weight_init()
        std = list(self.c.transition_lik_std)
        if len(std) == 4:
            std = std + 12 * [0.01]
        elif len(std) != cl // 2:
            raise ValueError('Specify valid transition_lik_std.')
        self.register_buffer('_transition_lik_std', torch.tensor([[std]], dtype=torch.float32),
                             persistent=False)
        self._perm_cache = {}

    @property
    def transition_lik_std(self):
        return self._transition_lik_std

    def weight_init(self):
        for mod in self.modules():
            if isinstance(mod, nn.Linear) and mod not in (getattr(self, 'action_embedding_layer', None),):
                nn.init.xavier_uniform_(mod.weight)
                nn.init.constant_(mod.bias, 0.1)

    # ------------------------------------------------------------------------------------
    def constrain_z_dyn(self, z, z_std=None):
        """dynamics.py:147-179: means to (-1, 1); stds to (0, pos_var | 0.04 | latent_q_std)."""
        z_c = 2 * torch.sigmoid(z) - 1
        if z_std is None:
            return z_c, None
        cache = self.__dict__.setdefault('_const_cache', {})
        key = (z_std.shape[-1], z_std.device, z_std.dtype)
        if key not in cache:
            cache[key] = torch.tensor([self.c.pos_var] * 2 + [0.04] * 2
                                      + [self.c.debug_latent_q_std] * (z_std.shape[-1] - 4),
                                      device=z_std.device, dtype=z_std.dtype)
        return z_c, cache[key] * torch.sigmoid(z_std)

    # -- fused path ----------------------------------------------------------------------
    def kernel_cfg(self, with_actions, with_app, lim_enc=2):
        nonlin = 1 if self.c.debug_nonlinear == 'leaky_relu' else 0     # inverted selector, dynamics.py:109
        return N.GnnCfg(self.c.num_obj, self.c.cl, self.c.action_space if with_actions else 0,
                        self.c.debug_appearance_dim if with_app else 0,
                        1 if self.c.action_conditioned else 0, lim_enc, nonlin,
                        0 if self.enc_input_size == self.c.cl // 2 else self.enc_input_size)

    def _segments(self, core_idx):
        """name -> (weight [out, in] or list of them to concatenate along out, bias)."""
        i = core_idx
        seg = {'enc': ([self.state_enc.weight], [self.state_enc.bias]),
               'self0': ([self.self_cores[i][0].weight], [self.self_cores[i][0].bias]),
               'self1': ([self.self_cores[i][1].weight], [self.self_cores[i][1].bias]),
               'ra0': ([self.rel_cores[i][0].weight, self.att_net[i][0].weight],
                       [self.rel_cores[i][0].bias, self.att_net[i][0].bias]),
               'rel1': ([self.rel_cores[i][1].weight], [self.rel_cores[i][1].bias]),
               'att1': ([self.att_net[i][1].weight], [self.att_net[i][1].bias]),
               'rel2': ([self.rel_cores[i][2].weight], [self.rel_cores[i][2].bias]),
               'att2': ([self.att_net[i][2].weight], [self.att_net[i][2].bias]),
               'aff0': ([self.affector[i][0].weight], [self.affector[i][0].bias]),
               'aff1': ([self.affector[i][1].weight], [self.affector[i][1].bias]),
               'aff2': ([self.affector[i][2].weight], [self.affector[i][2].bias]),
               'out0': ([self.out[i][0].weight], [self.out[i][0].bias]),
               'out1': ([self.out[i][1].weight], [self.out[i][1].bias])}
        if self.c.action_conditioned:
            seg['act'] = ([self.action_embedding_layer.weight], [self.action_embedding_layer.bias])
            h0, h1 = self.reward_head0, self.reward_head1
            seg['rew00'] = ([h0[0].weight], [h0[0].bias])
            seg['rew02'] = ([h0[2].weight], [h0[2].bias])
            seg['rew10'] = ([h1[0].weight], [h1[0].bias])
            seg['rew12'] = ([h1[2].weight], [h1[2].bias])
            seg['rew14'] = ([h1[4].weight], [h1[4].bias])
        return seg

    def _perm(self, cfg, core_idx, device):
        """Gather map raw-parameter-concat -> kernel layout ([in][out] matrices, padded)."""
        key = (cfg.action_dim, cfg.app_dim, cfg.lim_enc, cfg.state_dim, core_idx, str(device))
        if key in self._perm_cache:
            return self._perm_cache[key]
        off = ops.gnn_weight_offsets(cfg)
        seg = self._segments(core_idx)
        tensors, base, at = [], {}, 0
        for name, (ws, bs) in seg.items():
            for t in ws + bs:
                base[id(t)] = at
                tensors.append(t)
                at += t.numel()
        zero_idx = at
        perm = torch.full((off['total'],), zero_idx, dtype=torch.long)
        for name, (ws, bs) in seg.items():
            if off[name + '_w'] < 0:
                continue                              # e.g. 'act' when rolled out without actions
            n_total = sum(w.shape[0] for w in ws)
            col = 0
            for w in ws:
                o, i = w.shape
                k = torch.arange(i).view(i, 1)
                n = torch.arange(o).view(1, o)
                dst = off[name + '_w'] + k * n_total + (col + n)
                perm[dst.flatten()] = (base[id(w)] + n * i + k).flatten()
                col += o
            col = 0
            for b in bs:
                perm[off[name + '_b'] + col + torch.arange(b.numel())] = base[id(b)] + torch.arange(b.numel())
                col += b.numel()
        perm = perm.to(device)
        self._perm_cache[key] = (perm, [name for name in seg])
        return self._perm_cache[key]

    def pack_weights(self, core_idx=0, with_actions=None, with_app=None, lim_enc=2):
        """Flat kernel-layout weight buffer (differentiable w.r.t. the module parameters)."""
        with_actions = self.c.action_conditioned if with_actions is None else with_actions
        with_app = self.c.debug_core_appearance if with_app is None else with_app
        cfg = self.kernel_cfg(with_actions, with_app, lim_enc)
        in_dim = self.enc_input_size + (4 if with_actions else 0) + (self.c.debug_appearance_dim if with_app else 0)
        if in_dim != self.state_enc.in_features:
            raise ValueError('dynamics input has %d features but state_enc expects %d (actions / '
                             'appearances must match the configuration)' % (in_dim, self.state_enc.in_features))
        dev = self.state_enc.weight.device
        perm, _ = self._perm(cfg, core_idx, dev)
        seg = self._segments(core_idx)
        flat = [t.reshape(-1) for ws, bs in seg.values() for t in ws + bs]
        flat.append(torch.zeros(1, device=dev, dtype=self.state_enc.weight.dtype))
        return cfg, torch.cat(flat).index_select(0, perm)

    def forward(self, s, core_idx, actions=None, obj_appearances=None, lim_enc=2, packed=None):
        """s (n, O, enc_input_size = cl//2) -> (result (n, O, cl), reward (n, 1) | 0)   [dynamics.py:220-265]."""
        if self.c.action_conditioned and actions is None:
            raise ValueError('action-conditioned dynamics needs actions')
        if packed is None:
            packed = self.pack_weights(core_idx, actions is not None, obj_appearances is not None, lim_enc)
        cfg, weights = packed
        out, reward = ops.GnnStep.apply(s, actions, obj_appearances, weights, cfg)
        if self.c.action_conditioned:
            return out, reward
        return out, 0

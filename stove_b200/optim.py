"""Optimizer of the reference trainer on the flat gradient bucket.

The reference steps `optim.Adam(lr=c.learning_rate, amsgrad=c.debug_amsgrad)` after
`clip_grad_norm_(parameters, 1)` (model/video_prediction/train.py:46-49, 471-473) and anneals the
learning rate per step (`adjust_learning_rate`, train.py:154-159).  Here the global-norm clip and the
Adam(amsgrad) update are two launches on the data-parallel engine's flat bucket
(csrc/optim.cu); the step counter and the learning rate are device scalars, so the whole training
iteration -- forward, backward, all-reduce, clip, step -- can sit in one CUDA graph.
`state_dict()` / `load_state_dict()` use torch.optim.Adam's layout, so reference checkpoints
(train.py:78-116) round-trip.
"""
import ctypes as C
import math

import torch

from . import _native as N


class FusedAdam:
    def __init__(self, params, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, amsgrad=True, max_norm=1.0):
        self.all_params = list(params)                  # model.parameters() order = torch.optim index order
        self.lr, self.betas, self.eps, self.amsgrad = float(lr), tuple(betas), float(eps), bool(amsgrad)
        self.max_norm = max_norm
        self.live = None
        self.exp_avg = self.exp_avg_sq = self.max_exp_avg_sq = None
        self._table = None
        self._pending = None                            # optimizer state loaded before the first step

    # -- state ----------------------------------------------------------------------------------
    def _bind(self, live, flat):
        """First step: the live parameters (those that receive gradients) and their flat layout."""
        dev = flat.device
        self.live = list(live)
        numels = [p.numel() for p in self.live]
        offs, at = [], 0
        for m in numels:
            offs.append(at)
            at += m
        if at != flat.numel():
            raise ValueError('flat bucket holds %d floats, live parameters %d' % (flat.numel(), at))
        for p in self.live:
            if not p.is_contiguous():
                raise ValueError('parameters must be contiguous')
        N.require_cuda_f32(flat, *self.live)
        self.total = at
        n = len(self.live)
        self._table = ((C.c_void_p * n)(*[p.data_ptr() for p in self.live]), (C.c_int64 * n)(*offs),
                       (C.c_int64 * n)(*numels), n, offs)
        self.exp_avg = torch.zeros(at, device=dev)
        self.exp_avg_sq = torch.zeros(at, device=dev)
        self.max_exp_avg_sq = torch.zeros(at, device=dev) if self.amsgrad else None
        self.partial = torch.zeros(N.lib().stove_adam_workspace_floats(), device=dev)
        self.step_dev = torch.zeros(1, device=dev)
        self.lr_dev = torch.full((1,), self.lr, device=dev)
        if self._pending is not None:
            self._install(self._pending)
            self._pending = None

    def set_lr(self, lr):
        self.lr = float(lr)
        if self.live is not None:
            self.lr_dev.fill_(self.lr)

    def adjust_learning_rate(self, step, value, base_lr, min_lr):
        """train.py:154-159: lr = max(base * exp(-step / value), min_lr)."""
        self.set_lr(max(base_lr * math.exp(-step / value), min_lr))

    # -- step -----------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, live, flat):
        """`flat`: the engine's gradient bucket (already averaged over ranks), `live`: the parameters it
        covers, in bucket order.  Clips the bucket in place and updates the parameters in place."""
        if self.live is None:
            self._bind(live, flat)
        elif len(live) != len(self.live) or any(a is not b for a, b in zip(live, self.live)):
            raise RuntimeError('the set of parameters receiving gradients changed between steps')
        ptrs, offs, nums, n, _ = self._table
        b1, b2 = self.betas
        N.check(N.lib().stove_adam_step(C.cast(ptrs, C.c_void_p), C.cast(offs, C.c_void_p), C.cast(nums, C.c_void_p), n,
                                        self.total, N.ptr(flat), N.ptr(self.exp_avg), N.ptr(self.exp_avg_sq),
                                        N.ptr(self.max_exp_avg_sq), N.ptr(self.partial), N.ptr(self.lr_dev),
                                        N.ptr(self.step_dev), b1, b2, self.eps,
                                        float(self.max_norm) if self.max_norm else 0.0, N.stream()))

    # -- torch.optim.Adam-compatible checkpoints ------------------------------------------------------
    def state_dict(self):
        state = {}
        if self.live is not None:
            index = {id(p): i for i, p in enumerate(self.all_params)}
            step = self.step_dev.clone().cpu().reshape(())
            for p, off in zip(self.live, self._table[4]):
                sl = slice(off, off + p.numel())
                st = {'step': step.clone(), 'exp_avg': self.exp_avg[sl].view_as(p).clone(),
                      'exp_avg_sq': self.exp_avg_sq[sl].view_as(p).clone()}
                if self.amsgrad:
                    st['max_exp_avg_sq'] = self.max_exp_avg_sq[sl].view_as(p).clone()
                state[index[id(p)]] = st
        group = {'lr': self.lr, 'betas': self.betas, 'eps': self.eps, 'weight_decay': 0, 'amsgrad': self.amsgrad,
                 'maximize': False, 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None,
                 'params': list(range(len(self.all_params)))}
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        group = sd['param_groups'][0]
        self.betas, self.eps = tuple(group['betas']), float(group['eps'])
        if bool(group.get('amsgrad', False)) != self.amsgrad:
            raise ValueError('checkpoint amsgrad=%s, optimizer amsgrad=%s' % (group.get('amsgrad'), self.amsgrad))
        self.set_lr(group['lr'])
        if self.live is None:
            self._pending = sd['state']                 # installed when the live set is known (first step)
        else:
            self._install(sd['state'])

    def _install(self, state):
        index = {id(p): i for i, p in enumerate(self.all_params)}
        steps = set()
        for p, off in zip(self.live, self._table[4]):
            st = state.get(index[id(p)])
            if st is None:
                continue
            sl = slice(off, off + p.numel())
            self.exp_avg[sl].copy_(st['exp_avg'].reshape(-1))
            self.exp_avg_sq[sl].copy_(st['exp_avg_sq'].reshape(-1))
            if self.amsgrad:
                self.max_exp_avg_sq[sl].copy_(st['max_exp_avg_sq'].reshape(-1))
            steps.add(float(st['step']))
        if len(steps) > 1:
            raise ValueError('per-parameter step counts differ in the checkpoint: %s' % sorted(steps))
        if steps:
            self.step_dev.fill_(steps.pop())

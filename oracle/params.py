"""ORACLE (test infrastructure only): the reference's `state_dict()` layout and a
deterministic weight generator.

`param_shapes(c)` lists every key/shape of `Stove(config).state_dict()`
(model/video_prediction/stove.py:17-31, encoder.py:19-21, dynamics.py:29-98,
rat_torch.py:279-331 -- the root sum is registered twice, as `vector_list.L.0` and as
`output_vector`).  Checked against the live reference in
tests/test_oracle_vs_reference.py.

`make_state_dict(c, seed)` draws every tensor from a seeded CPU generator so the build
container and the GPU box construct bit-identical weights without shipping 6 MB of them.
"""
import math
from collections import OrderedDict

import torch

from .stove_oracle import structures


def param_shapes(c):
    obj_s, bg_s = structures(c)
    img = c.channels * c.width * c.height
    cl, O = c.cl, c.num_obj
    out = OrderedDict()
    out['sup.encoder.rnn.weight_ih_l0'] = (1024, img)
    out['sup.encoder.rnn.weight_hh_l0'] = (1024, 256)
    out['sup.encoder.rnn.bias_ih_l0'] = (1024,)
    out['sup.encoder.rnn.bias_hh_l0'] = (1024,)
    out['sup.encoder.fc1.weight'] = (50, 256)
    out['sup.encoder.fc1.bias'] = (50,)
    out['sup.encoder.fc2.weight'] = (8, 50)
    out['sup.encoder.fc2.bias'] = (8,)
    for name, s in (('obj_spn', obj_s), ('bg_spn', bg_s)):
        shapes = s.param_shapes()
        for k, v in shapes.items():
            out['sup.%s.%s' % (name, k)] = v
        root_key = 'vector_list.%d.%d.params' % s.root
        out['sup.%s.output_vector.params' % name] = shapes[root_key]
    enc_in = cl // 2
    if c.action_conditioned:
        out['dyn.action_embedding_layer.weight'] = (O * 4, c.action_space)
        out['dyn.action_embedding_layer.bias'] = (O * 4,)
        enc_in += 4
        for nm, (o, i) in (('reward_head0.0', (cl, cl)), ('reward_head0.2', (cl, cl)),
                           ('reward_head1.0', (cl // 2, cl)), ('reward_head1.2', (cl // 4, cl // 2)),
                           ('reward_head1.4', (1, cl // 4))):
            out['dyn.%s.weight' % nm] = (o, i)
            out['dyn.%s.bias' % nm] = (o,)
    if c.debug_core_appearance:
        enc_in += c.debug_appearance_dim
    out['dyn.state_enc.weight'] = (cl, enc_in)
    out['dyn.state_enc.bias'] = (cl,)
    groups = (('self_cores', [(cl, cl), (cl, cl)]),
              ('rel_cores', [(2 * cl, 2 * cl + 1), (cl, 2 * cl), (cl, cl)]),
              ('att_net', [(2 * cl, 2 * cl + 1), (cl, 2 * cl), (1, cl)]),
              ('affector', [(cl, cl), (cl, cl), (cl, cl)]),
              ('out', [(cl, 2 * cl), (cl, cl)]))
    for g, layers in groups:
        for core in range(3):
            for li, (o, i) in enumerate(layers):
                out['dyn.%s.%d.%d.weight' % (g, core, li)] = (o, i)
                out['dyn.%s.%d.%d.bias' % (g, core, li)] = (o,)
    return out


def make_state_dict(c, seed, dtype=torch.float64, att_gain=1.0):
    """Deterministic weights with init-like magnitudes (uniform +-1/sqrt(fan_in) for
    linear/LSTM tensors, clipped N(0, 0.1) for SPN tensors)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in param_shapes(c).items():
        if name.endswith('output_vector.params'):
            continue
        if '_spn.' in name:
            t = torch.randn(*shape, generator=g, dtype=torch.float64).clamp_(-2, 2) * 0.1
        else:
            if 'rnn' in name:
                bound = 1.0 / 16.0
            elif len(shape) == 2:
                bound = 1.0 / math.sqrt(shape[1])
            else:
                bound = 0.1
            t = (torch.rand(*shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
            if 'att_net' in name:
                t = t * att_gain
        sd[name] = t.to(dtype)
    for spn in ('obj_spn', 'bg_spn'):
        root = [k for k in sd if k.startswith('sup.%s.vector_list.' % spn) and k.endswith('.params')][-1]
        sd['sup.%s.output_vector.params' % spn] = sd[root]
    return sd

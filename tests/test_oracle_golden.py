"""CPU: the oracle (oracle/*.py) against the golden vectors produced by the unmodified
reference (oracle/make_golden.py).  This is what pins the oracle."""
import pytest
import torch

from oracle import stove_oracle as so
from oracle import spn_oracle as sp
from oracle.params import make_state_dict
from util import load_golden, load_structure_golden, rel_err, grad_signature, checksum, VARIANTS

D = torch.float64
TOL = 1e-10


def test_structure_matches_reference():
    gold = load_structure_golden()
    for tag, g in gold.items():
        s = sp.SpnStructure(g['n'], g['seed'], [tuple(x) for x in g['splits']], g['G'], g['S'])
        d = s.describe()
        assert d['leaf_scopes'] == g['leaf_scopes'], tag
        assert d['wiring'] == g['wiring'], tag


def _params(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items() if 'output_vector' not in k}


def test_spn_forward_backward():
    g = load_golden('spn')
    c = so.default_config()
    sd = make_state_dict(c, int(g['seed']))
    assert abs(checksum(sd) - float(g['checksum'])) < 1e-6
    obj_s, bg_s = so.structures(c)
    for tag, s, x, m, lo, hi in (('obj', obj_s, g['xo'], g['mo'], c.obj_min_var, c.obj_max_var),
                                 ('bg', bg_s, g['xb'], g['mb'], c.bg_min_var, c.bg_max_var)):
        P = _params(sd)
        pre = 'sup.%s_spn.' % tag
        xg, mg = x.clone().requires_grad_(True), m.clone().requires_grad_(True)
        out = sp.spn_forward(s, P, xg, mg, lo, hi, prefix=pre)
        assert rel_err(out, g[tag + '_out']) < TOL
        assert rel_err(sp.spn_forward(s, P, x, None, lo, hi, prefix=pre), g[tag + '_out_nomarg']) < TOL
        w = torch.linspace(0.5, 1.5, out.shape[0], dtype=D).unsqueeze(1)
        (out * w).sum().backward()
        assert rel_err(xg.grad, g[tag + '_gx']) < TOL
        assert rel_err(mg.grad, g[tag + '_gm']) < TOL
        for k in g:
            if k.startswith(tag + '_g.'):
                assert rel_err(P[pre + k[len(tag) + 3:]].grad, g[k]) < TOL, k


def test_scene_likelihood():
    g = load_golden('scene')
    c = so.default_config()
    sd = make_state_dict(c, int(g['seed']))
    P = _params(sd)
    x = so.bw_transform(g['x_u8'].to(D) / 255.0)
    z = g['z'].clone().requires_grad_(True)
    mp, mb, ov = so.masks_from_z(c, z)
    assert rel_err(mp, g['marg_patch']) < TOL
    assert rel_err(mb, g['marg_bg']) < TOL
    assert rel_err(ov, g['overlap']) < TOL
    assert rel_err(so.patches_from_z(c, x.flatten(0, 1), z.flatten(0, 1)), g['patches']) < TOL
    ll = so.likelihood(c, P, so.structures(c), x, z.flatten(0, 1))
    assert rel_err(ll, g['ll']) < TOL
    (ll * g['w']).sum().backward()
    assert rel_err(z.grad, g['gz']) < 1e-9
    for k in g:
        if k.startswith('g.sup.'):
            assert rel_err(P['sup.' + k[6:]].grad, g[k]) < 1e-9, k


@pytest.mark.parametrize('tag', ['plain', 'ac', 'o6'])
def test_dynamics(tag):
    g = load_golden('dynamics')
    kw = {'plain': {}, 'ac': VARIANTS['ac'][0], 'o6': dict(num_obj=6, debug_match_objects='greedy')}[tag]
    c = so.default_config(**kw)
    sd = make_state_dict(c, 13)
    assert abs(checksum(sd) - float(g[tag + '_checksum'])) < 1e-6
    P = _params(sd)
    s = g[tag + '_s'].clone().requires_grad_(True)
    a = g.get(tag + '_a')
    app = g.get(tag + '_app')
    out, rew = so.dynamics_forward(c, P, s, 0, a, app)
    assert rel_err(out, g[tag + '_out']) < TOL
    loss = (out * g[tag + '_w']).sum()
    if c.action_conditioned:
        assert rel_err(rew, g[tag + '_rew']) < TOL
        loss = loss + (rew * torch.linspace(1, 2, s.shape[0], dtype=D).unsqueeze(1)).sum()
    loss.backward()
    assert rel_err(s.grad, g[tag + '_gs']) < TOL
    for k in g:
        if k.startswith(tag + '_g.'):
            assert rel_err(P[k[len(tag) + 3:]].grad, g[k]) < TOL, k


@pytest.mark.parametrize('tag', list(VARIANTS))
def test_stove_forward_backward_rollout(tag):
    kw, seed = VARIANTS[tag]
    g = load_golden('stove_' + tag)
    c = so.default_config(**kw)
    sd = make_state_dict(c, seed, att_gain=float(g['att_gain']))
    assert abs(checksum(sd) - float(g['checksum'])) < 1e-6
    P = _params(sd)
    x = g['x_u8'].to(D) / 255.0
    n, T = x.shape[0], x.shape[1]
    noise = [g['noise%d' % i] for i in range(2 + T - c.skip)]
    actions = g.get('actions')
    parts = {}
    elbo, prop, rew = so.stove_forward(c, P, x, noise, actions=actions, parts=parts)
    assert rel_err(elbo, g['elbo']) < TOL
    for k in ('z', 'z_dyn', 'z_sup', 'log_q', 'translik'):
        assert rel_err(prop[k], g[k]) < TOL, k
    assert rel_err(parts['bg'].mean(), g['bg']) < TOL
    assert rel_err(parts['patch'].mean(), g['patch']) < TOL
    assert rel_err(parts['overlap'].mean(), g['overlap']) < TOL
    loss = -elbo
    if c.action_conditioned:
        assert rel_err(rew, g['rewards']) < TOL
        assert rel_err(prop['obj_appearances'], g['obj_appearances']) < TOL
        tgt = (torch.arange(n * (T - 2)) % 3 == 0).to(D).view(n, T - 2, 1)
        loss = loss + 100.0 * torch.nn.functional.binary_cross_entropy(rew, tgt)
    loss.backward()
    for k in g:
        if k.startswith('g.'):
            assert rel_err(P[k[2:]].grad, g[k]) < 1e-8, k
        elif k.startswith('gsig.'):
            sig = grad_signature(P[k[5:]].grad)
            assert float((sig - g[k]).abs().max() / g[k].abs().max()) < 1e-8, k
    with torch.no_grad():
        z_last = prop['z'][:, -1]
        steps = g['roll_steps'].tolist()
        if c.action_conditioned:
            zr, rr = so.rollout(c, P, z_last, num=92, actions=g['roll_actions'],
                                appearance=prop['obj_appearances'][:, -1])
            assert rel_err(rr[:, steps], g['roll_rewards']) < 1e-9
        else:
            zr, _ = so.rollout(c, P, z_last, num=92)
    assert rel_err(zr[:, steps], g['roll_z']) < 1e-9


def test_supair_only_elbo_golden():
    """Pretraining branch (Stove.forward(..., pretrain=True) -> Supair.forward, supair.py:504-551) against the
    reference-generated golden."""
    g = load_golden('supair_only')
    c = so.default_config()
    sd = make_state_dict(c, int(g['seed']))
    assert abs(checksum(sd) - float(g['checksum'])) < 1e-6
    P = _params(sd)
    x = g['x_u8'].to(D) / 255.0
    elbo, prop = so.supair_forward(c, P, x, [g['noise0']])
    assert rel_err(elbo, g['elbo']) < TOL
    assert rel_err(prop['z'], g['z']) < TOL
    (-elbo).backward()
    for k in g:
        if k.startswith('g.'):
            assert rel_err(P[k[2:]].grad, g[k]) < 1e-8, k
        elif k.startswith('gsig.'):
            sig = grad_signature(P[k[5:]].grad)
            assert float((sig - g[k]).abs().max() / g[k].abs().max()) < 1e-8, k

"""GPU parity: fused GNN step (forward + backward) vs the reference's golden vectors."""
import pytest
import torch

from oracle import stove_oracle as so
from util import Checker, VARIANTS, load_golden, make_model

pytestmark = pytest.mark.gpu
VAL, GRAD = 2e-5, 4e-4
KW = {'plain': {}, 'ac': VARIANTS['ac'][0], 'o6': dict(num_obj=6, debug_match_objects='greedy')}


@pytest.mark.parametrize('tag', ['plain', 'ac', 'o6'])
def test_dynamics_golden(tag):
    g = load_golden('dynamics')
    oc, sd, model = make_model(KW[tag], 13)
    s = g[tag + '_s'].float().cuda().requires_grad_(True)
    a = g[tag + '_a'].float().cuda() if tag + '_a' in g else None
    app = g[tag + '_app'].float().cuda() if tag + '_app' in g else None
    out, rew = model.dyn(s, 0, a, app)
    ck = Checker('dynamics_' + tag)
    ck.close('out', out, g[tag + '_out'], VAL)
    loss = (out * g[tag + '_w'].float().cuda()).sum()
    if oc.action_conditioned:
        ck.close('reward', rew, g[tag + '_rew'], VAL)
        loss = loss + (rew * torch.linspace(1, 2, s.shape[0], device='cuda').unsqueeze(1)).sum()
    else:
        ck.true('reward_is_zero', rew == 0)
    model.zero_grad()
    loss.backward()
    ck.close('gs', s.grad, g[tag + '_gs'], GRAD)
    params = dict(model.named_parameters())
    for k in g:
        if k.startswith(tag + '_g.'):
            ck.close(k, params[k[len(tag) + 3:]].grad, g[k], GRAD)
    # dead cores (1, 2) never receive gradients, like the reference (SURVEY hard part 10)
    ck.true('dead_core_grad_none', model.dyn.self_cores[1][0].weight.grad is None)
    ck.finish()


@pytest.mark.parametrize('n', [1, 5, 300, 1024])
def test_dynamics_batch_sizes_vs_oracle(n):
    oc, sd, model = make_model(VARIANTS['ac'][0], 14)
    gen = torch.Generator().manual_seed(n)
    s = torch.rand(n, 3, 16, generator=gen, dtype=torch.float64) * 1.6 - 0.8
    a = torch.nn.functional.one_hot(torch.randint(9, (n,), generator=gen), 9).double()
    app = torch.rand(n, 3, 3, generator=gen, dtype=torch.float64)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items() if 'output_vector' not in k}
    so_s = s.clone().requires_grad_(True)
    ro, rr = so.dynamics_forward(oc, P, so_s, 0, a, app)
    w = torch.sin(torch.arange(ro.numel(), dtype=torch.float64)).view_as(ro)
    ((ro * w).sum() + rr.sum()).backward()
    sg = s.float().cuda().requires_grad_(True)
    out, rew = model.dyn(sg, 0, a.float().cuda(), app.float().cuda())
    model.zero_grad()
    ((out * w.float().cuda()).sum() + rew.sum()).backward()
    ck = Checker('dynamics_n%d' % n)
    ck.close('out', out, ro, VAL)
    ck.close('reward', rew, rr, VAL)
    ck.close('gs', sg.grad, so_s.grad, GRAD)
    for name, p in model.dyn.named_parameters():
        if p.grad is not None:
            ck.close('g.' + name, p.grad, P['dyn.' + name].grad, GRAD)
    ck.finish()


@pytest.mark.parametrize('cl,enc,lim,ac', [(16, 16, 4, False), (16, 16, 4, True), (32, 16, 2, False), (24, 20, 4, False)])
def test_dynamics_parametric_shapes_vs_oracle(cl, enc, lim, ac):
    """Other widths through the single-step kernels: the supervised ablation builds
    Dynamics(config, enc_input_size=16) on cl = 16 and calls forward(..., lim_enc=4)
    (supairvised/dynamics.py:24-25, 75-77)."""
    from stove_b200 import Dynamics, StoveConfig
    kw = dict(cl=cl, num_obj=3, width=32, height=32)
    if ac:
        kw.update(action_conditioned=True, action_space=9, debug_core_appearance=True)
    else:
        kw.update(action_conditioned=False, action_space=None)
    torch.manual_seed(cl + enc)
    dyn = Dynamics(StoveConfig(**kw), enc_input_size=enc).cuda()
    oc = so.default_config(**kw)
    P = {'dyn.' + k: v.detach().double().cpu().clone().requires_grad_(True) for k, v in dyn.state_dict().items()}
    n = 77
    gen = torch.Generator().manual_seed(n)
    s = torch.rand(n, 3, enc, generator=gen, dtype=torch.float64) * 1.6 - 0.8
    a = torch.nn.functional.one_hot(torch.randint(9, (n,), generator=gen), 9).double() if ac else None
    app = torch.rand(n, 3, 3, generator=gen, dtype=torch.float64) if ac else None
    so_s = s.clone().requires_grad_(True)
    ro, rr = so.dynamics_forward(oc, P, so_s, 0, a, app, lim_enc=lim)
    w = torch.sin(torch.arange(ro.numel(), dtype=torch.float64)).view_as(ro)
    ((ro * w).sum() + (rr.sum() if ac else 0)).backward()
    sg = s.float().cuda().requires_grad_(True)
    out, rew = dyn(sg, 0, a.float().cuda() if ac else None, app.float().cuda() if ac else None, lim_enc=lim)
    ((out * w.float().cuda()).sum() + (rew.sum() if ac else 0)).backward()
    ck = Checker('dynamics_cl%d_enc%d' % (cl, enc))
    ck.close('out', out, ro, VAL)
    ck.close('gs', sg.grad, so_s.grad, GRAD)
    for name, p in dyn.named_parameters():
        if p.grad is not None:
            ck.close('g.' + name, p.grad, P['dyn.' + name].grad, GRAD)
    ck.finish()

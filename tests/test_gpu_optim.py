"""GPU parity of the fused clip + Adam(amsgrad) step (csrc/optim.cu) against torch.optim.Adam +
clip_grad_norm_, the pair the reference trainer uses (train.py:46-49, 471-473), and checkpoint round trips
in torch.optim's state_dict layout."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
SHAPES = [(1,), (3, 5), (300, 300), (7,), (50, 256), (33, 3), (1024, 64)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in SHAPES]


def _grads(seed, scale):
    g = torch.Generator().manual_seed(seed)
    return [(scale * torch.randn(*s, generator=g)).cuda() for s in SHAPES]


@pytest.mark.parametrize('amsgrad,max_norm,scale', [(True, 1.0, 1.0), (True, 1.0, 1e-4), (False, None, 0.3)])
def test_fused_adam_matches_torch(amsgrad, max_norm, scale):
    from stove_b200 import ops
    from stove_b200.optim import FusedAdam
    mine, ref = _params(0), _params(0)
    dead = torch.nn.Parameter(torch.ones(4, device='cuda'))             # never receives a gradient
    opt = FusedAdam(mine + [dead], lr=2e-3, amsgrad=amsgrad, max_norm=max_norm)
    topt = torch.optim.Adam(ref, lr=2e-3, amsgrad=amsgrad)
    for it in range(6):
        gs = _grads(10 + it, scale)
        flat = ops.gather_flat(gs)
        if it == 3:
            opt.set_lr(5e-4)
            for grp in topt.param_groups:
                grp['lr'] = 5e-4
        opt.step(mine, flat)
        for p, g in zip(ref, gs):
            p.grad = g.clone()
        if max_norm:
            torch.nn.utils.clip_grad_norm_(ref, max_norm)
        topt.step()
        # the bucket holds the clipped gradient afterwards, like .grad after clip_grad_norm_
        assert torch.allclose(flat, torch.cat([p.grad.reshape(-1) for p in ref]), rtol=2e-6, atol=1e-12)
        for a, b in zip(mine, ref):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (it, float((a - b).abs().max()))
    assert torch.equal(dead.detach(), torch.ones(4, device='cuda'))
    # checkpoint: torch.optim.Adam loads our state, and we load torch's
    sd = opt.state_dict()
    t2 = torch.optim.Adam([p for p in mine] + [dead], lr=1.0, amsgrad=amsgrad)
    t2.load_state_dict(copy.deepcopy(sd))
    assert t2.param_groups[0]['lr'] == pytest.approx(5e-4)
    for i, p in enumerate(mine):
        assert torch.allclose(t2.state[p]['exp_avg'], topt.state[ref[i]]['exp_avg'], rtol=1e-5, atol=1e-7)
        assert float(t2.state[p]['step']) == 6
    fresh = FusedAdam(mine + [dead], lr=1.0, amsgrad=amsgrad, max_norm=max_norm)
    fresh.load_state_dict(sd)
    gs = _grads(99, scale)
    fresh.step(mine, ops.gather_flat(gs))
    for p, g in zip(ref, gs):
        p.grad = g.clone()
    if max_norm:
        torch.nn.utils.clip_grad_norm_(ref, max_norm)
    topt.step()
    for a, b in zip(mine, ref):
        assert torch.allclose(a, b, rtol=3e-6, atol=3e-7)


def test_graphed_train_step_counts_steps_and_trains():
    """Whole iteration (forward, backward, bucket, clip, Adam) captured in one graph: the device step
    counter advances per replay and the ELBO improves on a fixed batch."""
    from stove_b200 import dp, synth
    from stove_b200.optim import FusedAdam
    from util import make_model
    _, _, model = make_model({}, 3, device='cuda:0')
    x = synth.billiards(16, 8, 3, res=32, seed=5)['x'].cuda()
    eng = dp.DataParallel(model)
    opt = FusedAdam(model.parameters(), lr=2e-3)
    step = dp.GraphedStep(eng, x, optimizer=opt)
    losses = [float(step(x)) for _ in range(12)]
    assert float(opt.step_dev) == 12
    assert all(l == l for l in losses) and losses[-1] < losses[0]
    assert set(opt.state_dict()['state']) == {i for i, p in enumerate(model.parameters()) if p.grad is not None}

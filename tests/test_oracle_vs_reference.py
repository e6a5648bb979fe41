"""CPU, build container only: the oracle against the live, unmodified reference
(skipped where /root/reference is not mounted, e.g. on the GPU box)."""
import pytest
import torch

from oracle import ref_harness as rh
from oracle import stove_oracle as so
from oracle.params import make_state_dict, param_shapes
from util import rel_err, VARIANTS

pytestmark = pytest.mark.skipif(not rh.available(), reason='reference not mounted')
D = torch.float64


@pytest.mark.parametrize('tag', list(VARIANTS))
def test_live_forward_backward(tag):
    kw, seed = VARIANTS[tag]
    c = so.default_config(**kw)
    sd = make_state_dict(c, seed + 50)
    ref = rh.build_reference(c, sd)
    assert list(ref.state_dict().keys()) == list(param_shapes(c).keys())
    assert [tuple(v.shape) for v in ref.state_dict().values()] == list(param_shapes(c).values())
    n, T, W = 3, 8, c.width
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, T, 3, W, W, generator=g, dtype=D) * (torch.rand(n, T, 3, W, W, generator=g) > 0.7)
    actions = None
    if c.action_conditioned:
        actions = torch.nn.functional.one_hot(torch.randint(9, (n, T), generator=g), 9).to(D)
    with rh.quiet(), rh.NoiseTape() as tape, rh.default_dtype(D):
        elbo_r, prop_r, rew_r = ref(x, 0, actions=actions)
        (-elbo_r).backward()
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items() if 'output_vector' not in k}
    elbo_o, prop_o, rew_o = so.stove_forward(c, P, x, tape.draws, actions=actions)
    (-elbo_o).backward()
    assert rel_err(elbo_o, elbo_r.detach()) < 1e-12
    assert rel_err(prop_o['z'], prop_r['z']) < 1e-12
    for name, p in ref.named_parameters():
        if p.grad is None:
            assert P[name].grad is None
        else:
            assert rel_err(P[name].grad, p.grad) < 1e-9, name


def test_live_fp32_close_to_fp64():
    """States the achievable fp32 tolerance of the *reference itself* (SURVEY hard part 4)."""
    c = so.default_config()
    sd = make_state_dict(c, 77)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(3, 8, 3, 32, 32, generator=g, dtype=D) * (torch.rand(3, 8, 3, 32, 32, generator=g) > 0.7)
    ref64 = rh.build_reference(c, sd)
    with rh.quiet(), rh.NoiseTape() as tape, rh.default_dtype(D):
        e64, _, _ = ref64(x, 0)
    ref32 = rh.build_reference(c, sd, dtype=torch.float32)
    with rh.quiet(), rh.NoiseTape(replay=tape.draws), rh.default_dtype(torch.float32):
        e32, _, _ = ref32(x.float(), 0)
    assert rel_err(e32, e64.detach()) < 1e-5


def test_live_supair_only_elbo():
    """The pretraining branch Stove.forward(..., pretrain=True) -> Supair.forward (supair.py:504-551)."""
    c = so.default_config()
    sd = make_state_dict(c, 91)
    ref = rh.build_reference(c, sd)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(3, 8, 3, 32, 32, generator=g, dtype=D) * (torch.rand(3, 8, 3, 32, 32, generator=g) > 0.7)
    with rh.quiet(), rh.NoiseTape() as tape, rh.default_dtype(D):
        elbo_r, prop_r, zero = ref(x, 0, pretrain=True)
        (-elbo_r).backward()
    assert zero == 0 and len(tape.draws) == 1
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items() if 'output_vector' not in k}
    elbo_o, prop_o = so.supair_forward(c, P, x, tape.draws)
    (-elbo_o).backward()
    assert rel_err(elbo_o, elbo_r.detach()) < 1e-12
    assert rel_err(prop_o['z'], prop_r['z']) < 1e-12
    for name, p in ref.named_parameters():
        if p.grad is None:
            assert P[name].grad is None, name
        else:
            assert rel_err(P[name].grad, p.grad) < 1e-9, name


@pytest.mark.parametrize('cl,enc,lim', [(16, 16, 4), (32, 16, 2)])
def test_live_dynamics_parametric_shapes(cl, enc, lim):
    """Dynamics(config, enc_input_size) with lim_enc: the shape of the supervised ablation
    (supairvised/dynamics.py:24-25, 75-77) through the unmodified reference class."""
    import sys
    rh._import()
    from model.video_prediction.dynamics import Dynamics as RefDynamics
    c = so.default_config(cl=cl)
    rc = rh.reference_config(c)
    with rh.default_dtype(D):
        torch.manual_seed(cl + enc)
        ref = RefDynamics(rc, enc).type(D)
    P = {'dyn.' + k: v.detach().clone() for k, v in ref.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    s = torch.rand(6, 3, enc, generator=g, dtype=D) * 1.6 - 0.8
    with rh.quiet(), rh.default_dtype(D):
        out_r, rew_r = ref(s, 0, lim_enc=lim)
    out_o, rew_o = so.dynamics_forward(c, P, s, 0, lim_enc=lim)
    assert rew_r == 0 and rew_o == 0
    assert rel_err(out_o, out_r.detach()) < 1e-12

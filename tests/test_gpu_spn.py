"""GPU parity: fused RAT-SPN kernels (through the C ABI) vs the reference's golden vectors
and vs the fp64 oracle.  Tolerances: fp32 kernels against fp64 gold, relative to the largest
reference magnitude of each tensor: 2e-5 for values, 2e-4 for gradients (the reference itself
in fp32 differs from fp64 by ~1e-6, SURVEY.md hard part 4)."""
import pytest
import torch

from oracle import spn_oracle as sp
from oracle import stove_oracle as so
from oracle.params import make_state_dict
from util import Checker, load_golden, make_model

pytestmark = pytest.mark.gpu
VAL, GRAD = 2e-5, 2e-4


def _spn(model, tag):
    return model.sup.obj_spn if tag == 'obj' else model.sup.bg_spn


def test_spn_golden_forward_backward():
    g = load_golden('spn')
    oc, sd, model = make_model({}, int(g['seed']))
    ck = Checker('spn_golden')
    for tag, xk, mk in (('obj', 'xo', 'mo'), ('bg', 'xb', 'mb')):
        spn = _spn(model, tag)
        x = g[xk].float().cuda().requires_grad_(True)
        m = g[mk].float().cuda().requires_grad_(True)
        out = spn(x, m)
        ck.close(tag + '_out', out, g[tag + '_out'], VAL)
        ck.close(tag + '_out_nomarg', spn(x.detach()), g[tag + '_out_nomarg'], VAL)
        w = torch.linspace(0.5, 1.5, out.shape[0], device='cuda').unsqueeze(1)
        model.zero_grad()
        (out * w).sum().backward()
        ck.close(tag + '_gx', x.grad, g[tag + '_gx'], GRAD)
        ck.close(tag + '_gm', m.grad, g[tag + '_gm'], GRAD)
        params = dict(spn.named_parameters())
        for k in g:
            if k.startswith(tag + '_g.'):
                ck.close(k, params[k[len(tag) + 3:]].grad, g[k], GRAD)
    ck.finish()


@pytest.mark.parametrize('tag,N', [('obj', 1), ('obj', 33), ('obj', 4608), ('bg', 1), ('bg', 130), ('bg', 1536)])
def test_spn_vs_oracle_sizes(tag, N):
    """Ragged / full config-1 sizes against the oracle on the same seeded inputs."""
    oc, sd, model = make_model({}, 31)
    spn = _spn(model, tag)
    obj_s, bg_s = so.structures(oc)
    struct, lo, hi, D = (obj_s, oc.obj_min_var, oc.obj_max_var, 100) if tag == 'obj' else \
        (bg_s, oc.bg_min_var, oc.bg_max_var, 1024)
    gen = torch.Generator().manual_seed(N)
    x = torch.rand(N, D, generator=gen, dtype=torch.float64)
    m = (torch.rand(N, D, generator=gen, dtype=torch.float64) * 1.4 - 0.2)
    m = torch.where(torch.rand(N, D, generator=gen) < 0.5, m.round().clamp(0, 1), m)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items() if 'output_vector' not in k}
    pre = 'sup.%s_spn.' % tag
    xo, mo = x.clone().requires_grad_(True), m.clone().requires_grad_(True)
    ref = sp.spn_forward(struct, P, xo, mo, lo, hi, prefix=pre)
    w = torch.cos(torch.arange(N, dtype=torch.float64)).unsqueeze(1) + 1.5
    (ref * w).sum().backward()
    xg, mg = x.float().cuda().requires_grad_(True), m.float().cuda().requires_grad_(True)
    out = spn(xg, mg)
    model.zero_grad()
    (out * w.float().cuda()).sum().backward()
    ck = Checker('spn_sizes_%s_%d' % (tag, N))
    ck.close('out', out, ref, VAL)
    # uniform-random frames give background leaf sums of magnitude ~5e3, whose fp32 ulp (5e-4)
    # bounds how well exp(val - logsumexp) can be known: allow 5e-4 on those gradients
    gtol = GRAD if tag == 'obj' else 5e-4
    ck.close('gx', xg.grad, xo.grad, gtol)
    ck.close('gm', mg.grad, mo.grad, gtol)
    for name, p in spn.named_parameters():
        if 'output_vector' not in name:
            ck.close('g.' + name, p.grad, P[pre + name].grad, gtol)
    ck.finish()


def test_spn_empty_batch_and_slow_path():
    oc, sd, model = make_model({}, 31)
    spn = model.sup.obj_spn
    assert spn(torch.zeros(0, 100, device='cuda')).shape == (0, 1)
    # sharply peaked sum weights + far-off inputs drive the linear-domain sums below
    # LIN_SUM_FLOOR, exercising the exact log-domain path (value and gradient)
    with torch.no_grad():
        for v in spn.vector_list[2]:
            v.params.mul_(0).add_(torch.linspace(-150, 150, 100, device='cuda').unsqueeze(1)
                                  * torch.linspace(0.2, 1.0, 10, device='cuda'))
        spn.output_vector.params.copy_(torch.linspace(-120, 120, 600, device='cuda').unsqueeze(1))
        for v in spn.vector_list[0]:
            v.means.copy_(torch.linspace(-3, 3, 10, device='cuda').expand_as(v.means))
    sd2 = {k: v.detach().double().cpu() for k, v in model.state_dict().items()}
    P = {k: v.clone().requires_grad_(True) for k, v in sd2.items() if 'output_vector' not in k}
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(40, 100, generator=gen, dtype=torch.float64) * 4 - 2
    xo = x.clone().requires_grad_(True)
    ref = sp.spn_forward(so.structures(oc)[0], P, xo, None, oc.obj_min_var, oc.obj_max_var,
                         prefix='sup.obj_spn.')
    ref.sum().backward()
    xg = x.float().cuda().requires_grad_(True)
    out = spn(xg)
    model.zero_grad()
    out.sum().backward()
    ck = Checker('spn_slow_path')
    ck.close('out', out, ref, 5e-5)
    ck.close('gx', xg.grad, xo.grad, 5e-4)
    # most root branches have responsibilities ~e^-100 here, so whole tensors have gradients
    # below the fp32 range: compare per parameter kind, relative to the largest gradient of that kind
    for kind in ('means', 'sigma_params', 'params'):
        names = [n for n, _ in spn.named_parameters() if n.endswith('.' + kind) and 'output_vector' not in n]
        got = torch.cat([dict(spn.named_parameters())[n].grad.flatten() for n in names])
        want = torch.cat([P['sup.obj_spn.' + n].grad.flatten() for n in names])
        ck.close('g.' + kind, got, want, 5e-4)
    ck.finish()

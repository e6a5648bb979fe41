"""ORACLE (test infrastructure only): generate tests/golden/*.npz by executing the
UNMODIFIED reference (/root/reference) in the build container, fp64.

    python -m oracle.make_golden            # rewrites tests/golden/

Weights are not stored: they are re-created from a seed by `oracle.params.make_state_dict`
(checksummed in every fixture).  Frames are stored as uint8 (k/255 exactly).
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_harness as rh                     # noqa: E402
from oracle import stove_oracle as so                    # noqa: E402
from oracle.params import make_state_dict                # noqa: E402
from stove_b200 import synth                             # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
D = torch.float64


def checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sd.items() if 'output_vector' not in k))


def grad_signature(g):
    """Compact, order-sensitive signature of a big gradient tensor."""
    flat = g.flatten().double()
    idx = torch.arange(flat.numel(), dtype=torch.float64)
    probe = torch.cos(idx * 0.37) + 0.5 * torch.sin(idx * 0.011)
    return np.array([flat.sum().item(), flat.abs().sum().item(), (flat * probe).sum().item(),
                     flat.norm().item()])


def frames_u8(n, T, O, res, seed, radius, **kw):
    x = synth.billiards(n, T, O, res=res, seed=seed, radius=radius, **kw)['x']
    q = torch.round(x * 255).to(torch.uint8)
    return q, q.to(D) / 255.0


def structure_golden():
    sys.path.insert(0, rh.REFERENCE_ROOT)
    from model.spn.region_graph import RegionGraph
    from model.spn.rat_torch import RatSpn, SpnArgs
    out = {}
    for tag, n, seed, splits, G, S in (('obj100_s7', 100, 7, [(2, 2)] * 6, 10, 10),
                                      ('obj100_s12345', 100, 12345, [(2, 2)] * 6, 10, 10),
                                      ('obj300_s7', 300, 7, [(2, 2)] * 6, 10, 10),
                                      ('bg1024_s7', 1024, 7, [(2, 1)] * 3, 6, 3),
                                      ('bg2500_s7', 2500, 7, [(2, 1)] * 3, 6, 3)):
        rg = RegionGraph(range(n), seed=seed)
        for parts, depth in splits:
            rg.random_split(parts, depth)
        args = SpnArgs()
        args.num_gauss, args.num_sums = G, S
        spn = RatSpn(1, region_graph=rg, args=args, name='g')
        index = {}
        for l, layer in enumerate(spn.vector_list):
            for i, v in enumerate(layer):
                index[v.name] = (l, i)
        scopes = [[int(p) for p in v.scope] for v in spn.vector_list[0]]
        wiring = [[[list(index[inp.name]) for inp in v.inputs] for v in layer]
                  for layer in list(spn.vector_list)[1:]]
        out[tag] = {'n': n, 'seed': seed, 'splits': splits, 'G': G, 'S': S,
                    'leaf_scopes': scopes, 'wiring': wiring}
    with open(os.path.join(OUT, 'spn_structure.json'), 'w') as f:
        json.dump(out, f)


def spn_golden():
    c = so.default_config()
    sd = make_state_dict(c, 11)
    ref = rh.build_reference(c, sd)
    g = torch.Generator().manual_seed(5)
    xo = torch.rand(9, 100, generator=g, dtype=D)
    mo = torch.rand(9, 100, generator=g, dtype=D) * 1.6 - 0.3      # exercises the clamp
    mo[0] = 0.0
    mo[1] = 1.0
    xb = torch.rand(5, 1024, generator=g, dtype=D)
    mb = (torch.rand(5, 1024, generator=g, dtype=D) > 0.6).to(D)
    mb[0] = torch.rand(1024, generator=g, dtype=D) * 1.4 - 0.2
    res = {'seed': 11, 'checksum': checksum(sd), 'xo': xo, 'mo': mo, 'xb': xb, 'mb': mb}
    for tag, spn, x, m in (('obj', ref.sup.obj_spn, xo, mo), ('bg', ref.sup.bg_spn, xb, mb)):
        xg = x.clone().requires_grad_(True)
        mg = m.clone().requires_grad_(True)
        out = spn.forward(xg, mg)
        res[tag + '_out'] = out.detach()
        res[tag + '_out_nomarg'] = spn.forward(x).detach()
        w = torch.linspace(0.5, 1.5, out.shape[0], dtype=D).unsqueeze(1)
        spn.zero_grad()
        (out * w).sum().backward()
        res[tag + '_gx'] = xg.grad
        res[tag + '_gm'] = mg.grad
        for name, p in spn.named_parameters():
            if 'output_vector' in name:
                continue
            res['%s_g.%s' % (tag, name)] = p.grad.clone()
    np.savez_compressed(os.path.join(OUT, 'spn.npz'), **{k: np.asarray(v) for k, v in res.items()})


def scene_golden():
    """Supair.likelihood and its pieces (supair.py:44-110, 241-356) on hand-made z that
    covers out-of-frame boxes, overlaps and tiny/huge scales."""
    c = so.default_config()
    sd = make_state_dict(c, 12)
    ref = rh.build_reference(c, sd)
    q, x = frames_u8(2, 4, 3, 32, 3, 1.2)
    xbw = so.bw_transform(x)                                   # (2, 4, 1, 32, 32)
    g = torch.Generator().manual_seed(6)
    F_ = 8
    z = torch.zeros(F_, 3, 4, dtype=D)
    z[..., 0] = 0.1 + 0.7 * torch.rand(F_, 3, generator=g, dtype=D)
    z[..., 1] = z[..., 0] * (0.75 + 0.5 * torch.rand(F_, 3, generator=g, dtype=D))
    z[..., 2:] = 0.9 * (2 * torch.rand(F_, 3, 2, generator=g, dtype=D) - 1)
    z[0, 0] = torch.tensor([0.8, 1.0, 0.85, -0.9])             # mostly out of frame
    z[0, 1] = torch.tensor([0.1, 0.075, 0.0, 0.0])             # tiny
    z[1, 1] = z[1, 0] + 0.03                                   # heavy overlap
    z[1, 2] = z[1, 0]
    zg = z.clone().requires_grad_(True)
    with rh.quiet(), rh.default_dtype(D):     # Exponential(beta) takes the default dtype
        ref.sup.step_counter = 1
        ref.zero_grad()
        marg_p, marg_bg, overlap = ref.sup.masks_from_z(zg)
        patches = ref.sup.patches_from_z(xbw.flatten(0, 1), zg.flatten(0, 1))
        ll, _ = ref.sup.likelihood(xbw, zg.flatten(0, 1))
        w = torch.linspace(0.7, 1.3, F_, dtype=D)
        (ll * w).sum().backward()
    res = {'seed': 12, 'checksum': checksum(sd), 'x_u8': q, 'z': z, 'patches': patches.detach(),
           'marg_patch': marg_p.detach(), 'marg_bg': marg_bg.detach(), 'overlap': overlap.detach(),
           'll': ll.detach(), 'gz': zg.grad, 'w': w}
    for name, p in ref.sup.named_parameters():
        if p.grad is not None and 'output_vector' not in name and 'spn' in name:
            res['g.sup.' + name] = p.grad.clone()
    np.savez_compressed(os.path.join(OUT, 'scene.npz'), **{k: np.asarray(v) for k, v in res.items()})


def dynamics_golden():
    res = {}
    for tag, kw in (('plain', {}),
                    ('ac', dict(action_conditioned=True, action_space=9, debug_core_appearance=True)),
                    ('o6', dict(num_obj=6, debug_match_objects='greedy'))):
        c = so.default_config(**kw)
        sd = make_state_dict(c, 13)
        ref = rh.build_reference(c, sd)
        g = torch.Generator().manual_seed(7)
        n, O = 6, c.num_obj
        s = (torch.rand(n, O, 16, generator=g, dtype=D) * 2 - 1) * 0.8
        a = app = None
        if c.action_conditioned:
            a = torch.nn.functional.one_hot(torch.randint(9, (n,), generator=g), 9).to(D)
            app = torch.rand(n, O, 3, generator=g, dtype=D)
        sg = s.clone().requires_grad_(True)
        ref.zero_grad()
        out, rew = ref.dyn(sg, 0, a, app)
        w = torch.cos(torch.arange(out.numel(), dtype=D) * 0.7).view_as(out)
        loss = (out * w).sum()
        if c.action_conditioned:
            loss = loss + (rew * torch.linspace(1, 2, n, dtype=D).unsqueeze(1)).sum()
        loss.backward()
        res.update({tag + '_s': s, tag + '_out': out.detach(), tag + '_gs': sg.grad,
                    tag + '_w': w, tag + '_checksum': checksum(sd)})
        if c.action_conditioned:
            res.update({tag + '_a': a, tag + '_app': app, tag + '_rew': rew.detach()})
        for name, p in ref.dyn.named_parameters():
            if p.grad is not None:
                res['%s_g.dyn.%s' % (tag, name)] = p.grad.clone()
    np.savez_compressed(os.path.join(OUT, 'dynamics.npz'), **{k: np.asarray(v) for k, v in res.items()})


def env_frames_u8(n, T, seed):
    """Frames from the reference's own simulator + renderer: BillardsEnv(n=3, r=1.2, m=1, hw=10, granularity=10,
    res=32, t=1, fc=0) through generate_fitting_run (model/envs/envs.py:588-653, 829-835),
    imported with the stubs of SURVEY appendix B (imageio / spriteworld are not needed by this path).
    -> (uint8 frames (n, T, 3, 32, 32), the same as float64 / 255)."""
    import sys
    import types
    for m in ['imageio', 'spriteworld', 'spriteworld.renderers', 'spriteworld.sprite']:
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['spriteworld'].renderers = sys.modules['spriteworld.renderers']
    sys.modules['spriteworld.sprite'].Sprite = object
    if rh.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, rh.REFERENCE_ROOT)
    from model.envs import envs
    envs.tqdm = lambda it, *a, **k: it                    # quiet
    np.random.seed(seed)
    imgs, _ = envs.generate_fitting_run(envs.BillardsEnv, run_len=T, run_num=n, max_tries=100 * n, res=32, n=3,
                                        r=1.2, dt=1, granularity=10, fc=0, hw=10, m=1.)
    assert imgs.shape == (n, T, 32, 32, 3), imgs.shape
    x = np.transpose(imgs, (0, 1, 4, 2, 3))                # the loader's (n, T, c, w, h) layout (load_data.py)
    q = np.clip(np.round(x * 255.0), 0, 255).astype(np.uint8)
    return torch.from_numpy(q), torch.from_numpy(q.astype(np.float64) / 255.0)


def stove_golden(only=None):
    variants = (
        ('plain', {}, 4, 32, 1.2, 21),
        ('ac', dict(action_conditioned=True, action_space=9, debug_core_appearance=True), 4, 32, 1.0, 22),
        ('o6', dict(num_obj=6, width=50, height=50, debug_match_objects='greedy', overlap_beta=100.0,
                    max_obj_scale=0.22), 2, 50, 1.0, 23),
        ('vol', dict(debug_match_objects='volatile'), 3, 32, 1.2, 24),
        ('envs', {}, 4, 32, 1.2, 26),          # frames rendered by the reference's own BillardsEnv
    )
    for tag, kw, n, res_px, radius, seed in variants:
        if only is not None and tag not in only:
            continue
        c = so.default_config(**kw)
        # att_gain < 1 keeps exp(attention) tame so long rollouts stay finite (SURVEY hard part 13)
        sd = make_state_dict(c, seed, att_gain=0.5)
        ref = rh.build_reference(c, sd)
        T = 8
        if tag == 'envs':
            q, x = env_frames_u8(n, T, seed)
        else:
            q, x = frames_u8(n, T, c.num_obj, res_px, seed, radius, use_colours=(c.num_obj <= 3))
        actions = None
        if c.action_conditioned:
            actions = synth.random_actions(n, T, 9, seed).to(D)
        torch.manual_seed(seed)
        with rh.quiet(), rh.NoiseTape() as tape, rh.default_dtype(D):
            ref.zero_grad()
            elbo, prop, rew = ref(x, 0, actions=actions)
            loss = -elbo
            if c.action_conditioned:
                tgt = (torch.arange(n * (T - 2)) % 3 == 0).to(D).view(n, T - 2, 1)
                loss = loss + 100.0 * torch.nn.functional.binary_cross_entropy(rew, tgt)
            loss.backward()
        out = {'seed': seed, 'att_gain': 0.5, 'checksum': checksum(sd), 'x_u8': q, 'elbo': elbo.detach(),
               'z': prop['z'], 'z_dyn': prop['z_dyn'], 'z_sup': prop['z_sup'],
               'log_q': prop['log_q'], 'translik': prop['translik'], 'bg': prop['bg'],
               'patch': prop['patch'], 'overlap': prop['overlap'],
               'rewards': torch.as_tensor(rew).detach()}
        for i, d in enumerate(tape.draws):
            out['noise%d' % i] = d
        if actions is not None:
            out['actions'] = actions
            out['obj_appearances'] = prop['obj_appearances']
        for name, p in ref.named_parameters():
            if p.grad is None:
                continue
            if p.numel() > 20000:
                out['gsig.' + name] = grad_signature(p.grad)
            else:
                out['g.' + name] = p.grad.clone()
        # rollouts seeded from the inferred state (stove.py:777-861)
        z_last = prop['z'][:, -1]
        app = prop['obj_appearances'][:, -1] if c.action_conditioned else None
        steps = [0, 1, 4, 9, 49, 91]
        with rh.quiet(), torch.no_grad():
            if c.action_conditioned:
                ract = synth.random_actions(n, 30, 9, seed + 100).to(D)   # wraps modulo 30
                zr, rr = ref.rollout(z_last, num=92, actions=ract, appearance=app)
                out['roll_actions'] = ract
                out['roll_rewards'] = rr[:, steps]
            else:
                zr, rr = ref.rollout(z_last, num=92)
        out['roll_steps'] = np.array(steps)
        out['roll_z'] = zr[:, steps]
        out['roll_finite'] = np.array(bool(torch.isfinite(zr).all()))
        np.savez_compressed(os.path.join(OUT, 'stove_%s.npz' % tag),
                            **{k: np.asarray(v) for k, v in out.items()})
        print(tag, 'elbo', float(elbo), 'rollout finite', bool(torch.isfinite(zr).all()),
              'max|pos|', float(zr[..., 2:4].abs().max()))


def supair_only_golden():
    """Pretraining branch: Stove.forward(x, 0, pretrain=True) -> Supair.forward (supair.py:504-551)."""
    seed = 25
    c = so.default_config()
    sd = make_state_dict(c, seed)
    ref = rh.build_reference(c, sd)
    n, T = 3, 8
    q, x = frames_u8(n, T, c.num_obj, 32, seed, 1.2)
    torch.manual_seed(seed)
    with rh.quiet(), rh.NoiseTape() as tape, rh.default_dtype(D):
        ref.zero_grad()
        elbo, prop, zero = ref(x, 0, pretrain=True)
        (-elbo).backward()
    assert zero == 0 and len(tape.draws) == 1
    out = {'seed': seed, 'checksum': checksum(sd), 'x_u8': q, 'elbo': elbo.detach(), 'z': prop['z'],
           'noise0': tape.draws[0]}
    for name, p in ref.named_parameters():
        if p.grad is None:
            continue
        if p.numel() > 20000:
            out['gsig.' + name] = grad_signature(p.grad)
        else:
            out['g.' + name] = p.grad.clone()
    np.savez_compressed(os.path.join(OUT, 'supair_only.npz'), **{k: np.asarray(v) for k, v in out.items()})
    print('supair_only elbo', float(elbo))


if __name__ == '__main__':
    assert rh.available(), 'needs /root/reference'
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1:                      # e.g. `python oracle/make_golden.py envs`: only these stove variants
        stove_golden(only=sys.argv[1:])
        sys.exit(0)
    structure_golden()
    spn_golden()
    scene_golden()
    dynamics_golden()
    stove_golden()
    supair_only_golden()
    print({f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})

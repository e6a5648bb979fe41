"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def load_structure_golden():
    with open(os.path.join(GOLDEN, 'spn_structure.json')) as f:
        return json.load(f)


def rel_err(a, b):
    """max |a-b| / max |b| (b = reference)."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def grad_signature(g):
    flat = g.detach().flatten().double().cpu()
    idx = torch.arange(flat.numel(), dtype=torch.float64)
    probe = torch.cos(idx * 0.37) + 0.5 * torch.sin(idx * 0.011)
    return torch.tensor([flat.sum().item(), flat.abs().sum().item(), (flat * probe).sum().item(),
                         flat.norm().item()], dtype=torch.float64)


def checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sd.items() if 'output_vector' not in k))


VARIANTS = {
    'plain': ({}, 21),
    'ac': (dict(action_conditioned=True, action_space=9, debug_core_appearance=True), 22),
    'o6': (dict(num_obj=6, width=50, height=50, debug_match_objects='greedy', overlap_beta=100.0,
                max_obj_scale=0.22), 23),
    'vol': (dict(debug_match_objects='volatile'), 24),
    'envs': ({}, 26),      # frames rendered by the reference's own BillardsEnv (oracle/make_golden.py: env_frames_u8)
}


class Checker:
    """Collects (name, error, tolerance) triples, logs them to gpurun_out/parity.jsonl and fails
    at the end, so one GPU round trip reports every mismatch instead of the first."""

    def __init__(self, test):
        self.test, self.rows = test, []

    def close(self, name, got, want, tol, absolute=False):
        got, want = got.detach().double().cpu(), want.detach().double().cpu()
        if got.shape != want.shape:
            self.rows.append((name, float('inf'), tol, 'shape %s vs %s' % (tuple(got.shape), tuple(want.shape))))
            return
        fin = torch.isfinite(want)
        bad_nan = bool((torch.isfinite(got) != fin).any())
        diff = (got - want)[fin].abs().max().item() if fin.any() else 0.0
        scale = 1.0 if absolute else (want[fin].abs().max().item() + 1e-300 if fin.any() else 1.0)
        err = float('inf') if bad_nan else diff / scale
        self.rows.append((name, err, tol, ''))

    def mostly_close(self, name, got, want, tol, max_bad_frac=1e-4):
        """Like close(), but a fraction of the entries may miss the tolerance: derivatives of bilinear sampling
        have kinks where a sample point crosses a pixel centre, and two fp32 evaluations of the coordinate can
        land on different sides (a handful of entries among millions of sample points)."""
        got, want = got.detach().double().cpu(), want.detach().double().cpu()
        if got.shape != want.shape:
            self.rows.append((name, float('inf'), tol, 'shape %s vs %s' % (tuple(got.shape), tuple(want.shape))))
            return
        if not bool(torch.isfinite(got).all()):
            self.rows.append((name, float('inf'), tol, 'non-finite'))
            return
        err = (got - want).abs() / (want.abs().max().item() + 1e-300)
        bad = int((err > tol).sum())
        allowed = int(max_bad_frac * err.numel())
        worst = float(err.max())
        self.rows.append((name, tol if bad <= allowed else worst, tol,
                          '%d of %d entries above tolerance (allowed %d), worst %.2e' % (bad, err.numel(), allowed, worst)))

    def true(self, name, cond):
        self.rows.append((name, 0.0 if cond else float('inf'), 0.5, ''))

    def finish(self):
        out = os.path.join(os.path.dirname(GOLDEN.rstrip('/')), '..', 'gpurun_out')
        try:
            os.makedirs(out, exist_ok=True)
            with open(os.path.join(out, 'parity.jsonl'), 'a') as f:
                for name, err, tol, note in self.rows:
                    f.write(json.dumps({'test': self.test, 'check': name, 'err': err, 'tol': tol,
                                        'note': note}) + '\n')
        except OSError:
            pass
        bad = [(n, e, t, note) for n, e, t, note in self.rows if not (e <= t)]
        worst = sorted(self.rows, key=lambda r: -(r[1] / r[2] if r[2] else 0))[:3]
        assert not bad, 'parity failures (name, err, tol): %s' % bad[:12]
        return worst


class NoiseReplay:
    """Stands in for Stove._standard_normal: replays a list of draws (shape-checked), one draw per
    call in the reference's order (`stacked = False`, see Stove._standard_normal_n).

    The draws stay referenced for the life of the replay object: the model consumes them on a SIDE stream (the
    packing stream, stove.py: stove_forward), and a draw that lost its last reference right after being handed
    out would return to the caching allocator -- which may give the block to the next allocation on the main
    stream before the side stream has read it.  (Found with compute-sanitizer racecheck slowing the GPU down:
    the golden test failed only in sequence, never with CUDA_LAUNCH_BLOCKING=1 or without the caching allocator.
    The product path draws its noise ON that side stream and is not affected.)"""
    stacked = False

    def __init__(self, draws, device):
        self.draws = [d.to(device=device, dtype=torch.float32) for d in draws]
        self.at = 0

    def __call__(self, shape, like):
        d = self.draws[self.at]
        self.at += 1
        assert tuple(d.shape) == tuple(shape), (d.shape, shape)
        return d


def make_model(kw, seed, att_gain=1.0, device='cuda'):
    """stove_b200.Stove with the deterministic weights of oracle.params.make_state_dict."""
    from oracle import stove_oracle as so
    from oracle.params import make_state_dict
    from stove_b200 import Stove, StoveConfig
    oc = so.default_config(**kw)
    sd = make_state_dict(oc, seed, att_gain=att_gain)
    cfg = StoveConfig(**{k: v for k, v in vars(oc).items()})
    model = Stove(cfg)
    model.load_state_dict({k: v.float() for k, v in sd.items()})
    return oc, sd, model.to(device)

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
}

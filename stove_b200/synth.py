"""Seeded synthetic video generators (billiards / gravity / avoidance-style actions).

The reference renders its data with numpy simulators (model/envs/envs.py:310-339 renderer,
:366-442 billiards, :445-524 gravity); they are not part of the hot path and need
packages absent here, so benchmarks and tests use this small stand-in that produces frames
of the same kind: O soft blobs `exp(-((d/r)^2)^4)` in separate colour channels on black,
float in [0, 1], shape (n, T, 3, res, res).  Physics is a plain elastic-collision /
pairwise-gravity integrator -- only the pixel statistics matter for throughput and parity.
"""
import numpy as np
import torch

_COLOURS = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1]], dtype=np.float64)


def _render(pos, radius, hw, res, use_colours):
    """pos (n, T, O, 2) in env units -> frames (n, T, 3, res, res)."""
    n, T, O, _ = pos.shape
    grid = (np.arange(res) + 0.5) * (hw / res)
    gi, gj = np.meshgrid(grid, grid)                       # gi varies along the last axis
    d2 = (gi[None, None, None] - pos[..., 0, None, None]) ** 2 + \
         (gj[None, None, None] - pos[..., 1, None, None]) ** 2
    blob = np.exp(-((d2 / radius ** 2) ** 4))              # (n, T, O, res, res)
    img = np.zeros((n, T, 3, res, res))
    for o in range(O):
        if use_colours:
            col = _COLOURS[o % len(_COLOURS)]
        else:
            col = np.eye(3)[o % 3]
        img += col[None, None, :, None, None] * blob[:, :, o, None]
    return np.minimum(img, 1.0)


def billiards(n, T=8, num_obj=3, res=32, hw=10.0, radius=1.2, seed=0, gravity=False,
              substeps=10, use_colours=True, burn_in=0):
    """Returns dict(x=(n,T,3,res,res) float32 frames, pos=(n,T,O,2), vel=(n,T,O,2))."""
    rng = np.random.RandomState(seed)
    O = num_obj
    # rejection-sample non-overlapping starts
    pos = np.zeros((n, O, 2))
    for o in range(O):
        todo = np.ones(n, dtype=bool)
        while todo.any():
            cand = radius + rng.rand(n, 2) * (hw - 2 * radius)
            ok = np.ones(n, dtype=bool)
            for p in range(o):
                ok &= np.linalg.norm(cand - pos[:, p], axis=1) > 2 * radius
            upd = todo & ok
            pos[upd, o] = cand[upd]
            todo &= ~ok
    vel = rng.randn(n, O, 2)
    vel = vel / np.linalg.norm(vel, axis=-1, keepdims=True) * (0.5 if not gravity else 0.3)
    dt = 1.0 / substeps
    out_p, out_v = [], []
    for t in range(burn_in + T):
        for _ in range(substeps):
            if gravity:
                diff = pos[:, None] - pos[:, :, None]                  # [b, i, j] = x_j - x_i
                dist = np.linalg.norm(diff, axis=-1) + 1e-9
                force = diff / np.maximum(dist, 2 * radius)[..., None] ** 3
                force[:, np.arange(O), np.arange(O)] = 0
                vel = vel + dt * 2.0 * force.sum(2)
            nxt = pos + vel * dt
            lo, hi = nxt < radius, nxt > hw - radius
            vel = np.where(lo | hi, -vel, vel)
            if not gravity:
                for i in range(O):
                    for j in range(i):
                        w = pos[:, i] - pos[:, j]
                        close = np.linalg.norm((pos[:, i] + vel[:, i] * dt) - (pos[:, j] + vel[:, j] * dt),
                                               axis=1) < 2 * radius
                        w = w / (np.linalg.norm(w, axis=1, keepdims=True) + 1e-9)
                        vi = (w * vel[:, i]).sum(1)
                        vj = (w * vel[:, j]).sum(1)
                        swap = (close[:, None] * w) * (vj - vi)[:, None]     # equal masses
                        vel[:, i] += swap
                        vel[:, j] -= swap
            pos = np.clip(pos + vel * dt, radius, hw - radius)
        if t >= burn_in:
            out_p.append(pos.copy())
            out_v.append(vel.copy())
    pos_t = np.stack(out_p, 1)
    x = _render(pos_t, radius, hw, res, use_colours)
    return {'x': torch.from_numpy(x.astype(np.float32)),
            'pos': torch.from_numpy(pos_t.astype(np.float32)),
            'vel': torch.from_numpy(np.stack(out_v, 1).astype(np.float32))}


def random_actions(n, T, action_space=9, seed=0):
    """One-hot float actions (n, T, A), like the MCTS driver's random policy
    (model/mcts/mcts_stove.py:118-125)."""
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(action_space, (n, T), generator=g)
    return torch.nn.functional.one_hot(idx, action_space).float()

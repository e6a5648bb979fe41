"""Batched MCTS leaf evaluation on the world model.

The reference's `BatchedMCTSHandler.run_mcts` (model/mcts/mcts_stove.py:94-137) evaluates the selected leaf
of every tree with two `Stove.rollout` calls per expansion round, each preceded by host-side one-hot
construction and `.to('cuda')` copies:
    1. expansion: every leaf state is tiled `action_space` times and advanced ONE step, one action each;
    2. value estimate: the expanded states are rolled out `2 * max_rollout_depth` steps under random actions.
The second call starts from the last state of the first, so the pair is one rollout of `1 + depth` steps whose
first action is the expansion action: `expand_and_rollout` issues it as ONE launch of the persistent rollout
kernel (csrc/dynloop.cu `team_rollout`), with the actions built on the device.  The tree logic (selection,
back-propagation) is orchestration and stays with the caller.
"""
import torch


def tile(a, dim, n_tile):
    """`tile` of the reference (mcts_stove.py:29-44): each slice along `dim` repeated n_tile times, in place."""
    return a.repeat_interleave(n_tile, dim=dim)


@torch.no_grad()
def expand_and_rollout(model, leaf_z, appearance, action_space=9, depth=40, generator=None, rollout_actions=None):
    """leaf_z (B, O, cl//2 + 2): the states selected in the B trees; appearance (B, O, 3) | None.

    Returns (new_zs (B*A, 1, O, Z), r (B*A, 1, 1), r_rollout (B*A, depth, 1)) exactly as the reference's two
    calls do: row j*A + a is leaf j expanded with action a.  `rollout_actions` (B*A, depth) int64 overrides the
    random policy (the reference draws them with numpy on the host, mcts_stove.py:118-125)."""
    B = leaf_z.shape[0]
    A = action_space
    dev = leaf_z.device
    z0 = tile(leaf_z, 0, A)
    app = tile(appearance, 0, A) if appearance is not None else None
    first = torch.arange(A, device=dev).repeat(B).unsqueeze(1)                       # (B*A, 1): 0 .. A-1 per leaf
    if rollout_actions is None:
        rollout_actions = torch.randint(A, (B * A, depth), device=dev, generator=generator)
    idx = torch.cat([first, rollout_actions.to(dev)], 1)                             # (B*A, 1 + depth)
    actions = torch.nn.functional.one_hot(idx, A).to(leaf_z.dtype)
    z_full, rewards = model.rollout(z0, num=1 + depth, actions=actions, appearance=app)
    return z_full[:, :1], rewards[:, :1], rewards[:, 1:]

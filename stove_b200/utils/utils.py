"""Helpers on the hot path (model/utils/utils.py:10-15)."""
from .. import ops


def bw_transform(x):
    """RGB-separated balls -> one channel: clamp(sum_c x, 0, 1), shape (n, T, 1, W, H)."""
    return ops.bw_transform(x)

"""Object and background SPN builders (drop-in for model/spn/probabilistic_models.py:8-39)."""
from .rat_torch import SpnArgs, RatSpn
from .region_graph import RegionGraph


def _get_obj_spn(c, seed):
    """Six `random_split(2, 2)` over the C*pw*ph patch pixels -> the "D2" fused structure."""
    rg = RegionGraph(range(c.channels * c.patch_width * c.patch_height), seed=seed)
    for _ in range(6):
        rg.random_split(2, 2)
    args = SpnArgs()
    args.num_gauss, args.num_sums = c.obj_spn_num_gauss, c.obj_spn_num_sums
    args.gauss_min_sigma, args.gauss_max_sigma = c.obj_min_var, c.obj_max_var
    return RatSpn(1, region_graph=rg, args=args, name='obj-spn')


def _get_bg_spn(c, seed):
    """Three `random_split(2, 1)` over all C*W*H pixels -> the "D1" fused structure."""
    rg = RegionGraph(range(c.width * c.height * c.channels), seed=seed)
    for _ in range(3):
        rg.random_split(2, 1)
    args = SpnArgs()
    args.num_gauss, args.num_sums = 6, 3
    args.gauss_min_sigma, args.gauss_max_sigma = c.bg_min_var, c.bg_max_var
    return RatSpn(1, region_graph=rg, args=args, name='bg-spn')

"""ORACLE (test infrastructure only): run the *unmodified* reference from /root/reference.

Reads /root/reference in the build container, or the staged copy oracle/_ref/ (oracle/build_ref.py;
git-ignored, shipped by gpurun) on the GPU box; used by `oracle/make_golden.py`,
`tests/test_oracle_vs_reference.py` (skipped when neither is there) and the reference arm of `bench.py`.  Shims are those of SURVEY.md appendix B: hand-built config,
distribution argument validation off (torch-1.0.1 behaviour), no `model.main` import.
"""
import contextlib
import os
import sys
import warnings

import torch

def _find_root():
    """/root/reference in the build container; the staged copy oracle/_ref (oracle/build_ref.py) on the GPU box"""
    for root in ('/root/reference', os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')):
        if os.path.isdir(os.path.join(root, 'model', 'video_prediction')):
            return root
    return '/root/reference'


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'model', 'video_prediction'))


def _import():
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    torch.distributions.Distribution.set_default_validate_args(False)
    from model.video_prediction.config import StoveConfig
    from model.video_prediction.stove import Stove
    return StoveConfig, Stove


def reference_config(oc, dtype=torch.float64):
    """Reference StoveConfig carrying the fields of an oracle config."""
    StoveConfig, _ = _import()
    c = StoveConfig()
    for k, v in vars(oc).items():
        if k != 'align_corners':
            setattr(c, k, v)
    c.device = torch.device('cpu')
    c.dtype = dtype
    c.num_frames, c.r, c.coord_lim = 100, 1.2, 10
    return c


@contextlib.contextmanager
def default_dtype(dtype):
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        yield
    finally:
        torch.set_default_dtype(old)


def build_reference(oc, state_dict=None, dtype=torch.float64):
    _, Stove = _import()
    with default_dtype(dtype):
        m = Stove(reference_config(oc, dtype)).type(dtype)
    if state_dict is not None:
        m.load_state_dict({k: v.to(dtype) for k, v in state_dict.items()})
    return m


class NoiseTape:
    """Record (or replay) the standard-normal draws behind every `Normal.rsample()`
    (torch.distributions.utils._standard_normal)."""

    def __init__(self, replay=None):
        self.draws = []
        self.replay = list(replay) if replay is not None else None

    def __enter__(self):
        import torch.distributions.normal as tdn
        self._mod, self._orig = tdn, tdn._standard_normal

        def fake(shape, dtype, device):
            if self.replay is not None:
                t = self.replay.pop(0).to(dtype)
                assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            else:
                t = self._orig(shape, dtype=dtype, device=device)
            self.draws.append(t.clone())
            return t

        tdn._standard_normal = fake
        return self

    def __exit__(self, *a):
        self._mod._standard_normal = self._orig


@contextlib.contextmanager
def quiet():
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        yield

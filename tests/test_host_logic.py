"""CPU: host-side logic of the package (no kernel launches): C-ABI exports, structure
tables, state_dict layout, sync-free object matching against the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import stove_oracle as so
from oracle.params import make_state_dict, param_shapes
from stove_b200 import Stove, StoveConfig, _native
from util import VARIANTS, load_structure_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'stove_b200.h')).read()
    declared = set(re.findall(r'^(?:int|int64_t|size_t|const char\*)\s+(stove_[a-z0-9_]+)\s*\(', header, re.M))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert _native.lib().stove_abi_version() == 1


def test_no_cpu_fallback():
    c = StoveConfig(width=32, height=32, num_obj=3, action_conditioned=False, random_seed=7)
    m = Stove(c)
    with pytest.raises(RuntimeError, match='CUDA'):
        m(torch.rand(2, 8, 3, 32, 32), 0)
    with pytest.raises(RuntimeError, match='CUDA'):
        m.sup.obj_spn(torch.rand(2, 100))


def test_product_code_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'stove_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f


@pytest.mark.parametrize('tag', list(VARIANTS))
def test_state_dict_layout_matches_reference(tag):
    kw, seed = VARIANTS[tag]
    oc = so.default_config(**kw)
    m = Stove(StoveConfig(**vars(oc)))
    want = param_shapes(oc)
    got = m.state_dict()
    assert list(got.keys()) == list(want.keys())
    assert [tuple(v.shape) for v in got.values()] == list(want.values())
    m.load_state_dict({k: v.float() for k, v in make_state_dict(oc, seed).items()})
    # the root sum is reachable under both names (rat_torch.py:331)
    assert m.sup.obj_spn.output_vector is m.sup.obj_spn.vector_list[4][0]


def test_structure_tables():
    from stove_b200.spn.rat_torch import RatSpn, SpnArgs
    from stove_b200.spn.region_graph import RegionGraph
    for tag, g in load_structure_golden().items():
        rg = RegionGraph(range(g['n']), seed=g['seed'])
        for p, d in g['splits']:
            rg.random_split(p, d)
        a = SpnArgs()
        a.num_gauss, a.num_sums = g['G'], g['S']
        spn = RatSpn(1, rg, a, name='t')
        assert [v.scope for v in spn.vector_list[0]] == g['leaf_scopes']
        t = spn._tables
        if tag.startswith('obj'):
            assert t.kind == 'D2'
            h = t.host
            D, R = g['n'], 6
            for p in range(D):
                for r in range(R):
                    q, pos = divmod(int(h['pix_slot'][p, r]), t.meta['pmax'])
                    assert q // 2 == r and h['region_scope'][q, pos] == p
            assert sorted(h['dst_row'].tolist()) == sorted(set(h['dst_row'].tolist()))
        else:
            assert t.kind == 'D1'
            assert t.host['side'].sum(0).tolist() == [g['n'] // 2] * 3
            # tables of the fused scene-likelihood kernels (csrc/scene_ll*.cu): scope lists per background leaf and the
            # row maps of the two lane-interleaved copies of the leaf table
            h, D, R = t.host, g['n'], 3
            for r in range(R):
                for side in (0, 1):
                    l, cnt = 2 * r + side, int(h['bg_cnt'][2 * r + side])
                    px = h['bg_scope'][l, :cnt]
                    assert (px[1:] > px[:-1]).all() and (h['side'][px, r] == side).all()
                    rows = h['il_f'][l * t.il_stride_f:l * t.il_stride_f + cnt]
                    assert (rows == px * R + r).all()                       # forward copy: (leaf, position in scope)
                    assert (h['il_f'][l * t.il_stride_f + cnt:(l + 1) * t.il_stride_f] == -1).all()
                assert int(h['bg_cnt'][2 * r]) + int(h['bg_cnt'][2 * r + 1]) == D
                rows = h['il_b'][r * t.il_stride_b:r * t.il_stride_b + D]
                assert (rows == np.arange(D) * R + r).all()                 # backward copy: (repetition, pixel)
            assert t.il_stride_f % 32 == 0 and t.il_stride_b % 32 == 0


@pytest.mark.parametrize('kind,O', [('3_only', 3), ('greedy', 6), ('volatile', 4), ('greedy', 3)])
def test_sync_free_matching_equals_oracle(kind, O):
    g = torch.Generator().manual_seed(O)
    oc = so.default_config(num_obj=O if kind != '3_only' else 3, debug_match_objects=kind)
    O = oc.num_obj
    m = Stove(StoveConfig(**vars(oc), width=32, height=32) if False else StoveConfig(**vars(oc)))
    n, T = 64, 8
    z = torch.rand(n, T, O, 4, generator=g) * 2 - 1
    z[:8, 3] = z[:8, 3, :1]              # force ambiguous / non-permutation argmins
    z[8:16, 5, 1] = z[8:16, 5, 0]
    std = torch.rand(n, T, O, 4, generator=g)
    for app in (None, torch.rand(n, T, O, 3, generator=g)):
        want = so.MATCHERS[kind](oc, z.double(), std.double(), app.double() if app is not None else None)
        got = m.match_objects(z, std, app)
        for a, b in zip(got, want):
            if b is None:
                assert a is None
            else:
                assert (a.double() - b).abs().max() < 1e-6


def test_fix_supair_and_velocities_equal_oracle():
    g = torch.Generator().manual_seed(3)
    oc = so.default_config()
    m = Stove(StoveConfig(**vars(oc)))
    z = torch.rand(16, 8, 3, 4, generator=g, dtype=torch.float64)
    z[:, 4] += 0.5
    std = torch.rand(16, 8, 3, 4, generator=g, dtype=torch.float64)
    a, b = m.fix_supair(z, std)
    a2, b2 = so.fix_supair(z, std)
    assert torch.equal(a, a2) and torch.equal(b, b2)
    assert torch.equal(m.v_from_state(z), so.v_from_state(z))
    assert torch.equal(m.v_std_from_pos(std), so.v_std_from_pos(std))
    zp = torch.randn(50, 8, generator=g, dtype=torch.float64)
    for x, y in zip(m.sup.constrain_zp(zp), so.constrain_zp(oc, zp)):
        assert (x - y).abs().max() < 1e-12


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the reference algorithm on the host cores): exactly one line on stdout, JSON,
    with the keys of the measurement contract."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'train_seqs_per_sec' and d['unit'] == 'sequences/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['n_gpus'] == 1
    from oracle import ref_harness as rh
    assert d['cpu_baseline']['kind'] == ('reference' if rh.available() else 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_shard_requires_an_even_split():
    """dp.shard: equal shards only (the exchange averages per-rank mean gradients with equal weights)."""
    import torch
    from stove_b200 import dp
    t = torch.arange(12).view(6, 2)
    assert torch.equal(dp.shard(t), t)                     # single process: the whole batch
    import unittest.mock as mock
    with mock.patch.object(dp, 'world', return_value=4), mock.patch.object(dp, 'rank', return_value=1):
        with pytest.raises(ValueError):
            dp.shard(t)
    with mock.patch.object(dp, 'world', return_value=3), mock.patch.object(dp, 'rank', return_value=2):
        assert torch.equal(dp.shard(t), t[4:6])


def test_split_k_heuristic():
    """split-K of the tensor-core GEMMs: about one CTA per SM, at least four k-blocks per part, capped."""
    from stove_b200.ops import _split_k
    assert _split_k(32, 32) == 4            # hidden-state gradient: 16 x 2 tiles, K = 1024
    assert _split_k(64, 64) == 2            # W_ih gradient: 8 x 8 tiles, K = 2048
    assert _split_k(16, 128, cap=4) == 4    # W_hh gradient
    assert _split_k(1, 1) == 1 and _split_k(200, 64) == 1


def test_reference_staging_recipe():
    """oracle/build_ref.py stages the unmodified modules of the path (git-ignored) where the reference is mounted."""
    from oracle import build_ref, ref_harness as rh
    dst = build_ref.build()
    if not os.path.isdir('/root/reference'):
        pytest.skip('reference not mounted')
    for rel in build_ref.FILES:
        src = os.path.join('/root/reference', rel)
        if os.path.exists(src):
            with open(src, 'rb') as a, open(os.path.join(dst, rel), 'rb') as b:
                assert a.read() == b.read(), rel            # byte-identical: nothing is edited on the way
    assert rh.available()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, '.gitignore')) as f:
        assert 'oracle/_ref/' in f.read()


def test_unsupported_debug_flags_raise():
    """Flags whose semantics the fused kernels do not implement must not silently compute the default."""
    from stove_b200 import Stove, StoveConfig
    for flag in ('debug_no_latents', 'debug_no_velocity'):
        cfg = StoveConfig(width=32, height=32, num_obj=3, action_conditioned=False, action_space=None, **{flag: True})
        with pytest.raises(NotImplementedError):
            Stove(cfg)


def test_scene_ll_planner_without_a_gpu():
    """stove_scene_ll_supported only plans (frames per round, tiles, shared memory): callable on the CPU.  The fused
    kernels take single-channel frames with the SuPAIR SPN sizes and give way to the unfused kernels otherwise."""
    import ctypes as C
    lib = _native.lib()
    obj = _native.Spn2Struct(100, 6, 10, 10, 50, None, None, None, None)
    bg32 = _native.Spn1Struct(32 * 32, 3, 6, None)
    bg50 = _native.Spn1Struct(50 * 50, 3, 6, None)
    sup = lambda F, O, Cc, A, B, o, b: lib.stove_scene_ll_supported(F, O, Cc, A, B, 10, 10, C.byref(o), C.byref(b))
    assert sup(1792, 3, 1, 32, 32, obj, bg32) == 1                  # BASELINE config 1
    assert sup(3584, 3, 1, 32, 32, obj, bg32) == 1                  # config 3 (two rounds per CTA)
    assert sup(1792, 9, 1, 50, 50, obj, bg50) == 1                  # config 4
    assert sup(1, 3, 1, 32, 32, obj, bg32) == 1
    assert sup(1792, 3, 3, 32, 32, obj, bg32) == 0                  # colour glimpses (object_embedding): unfused kernels
    assert sup(1792, 3, 1, 32, 32, _native.Spn2Struct(100, 6, 8, 8, 50, None, None, None, None), bg32) == 0
    assert sup(1792, 3, 1, 32, 32, obj, bg50) == 0                  # background SPN of another frame size
    big = _native.Spn1Struct(200 * 200, 3, 6, None)
    assert sup(64, 3, 1, 200, 200, obj, big) == 0                   # a frame does not fit shared memory


def test_schedule_switches_return_the_previous_setting():
    """The A/B switches of the step schedule (bench.py --no-scene-seq / --no-dyn-stream) are plain host flags: each setter
    returns the previous value, so a test can restore it.  Defaults: sequence-mode scene likelihood, fused backward and the
    dynamics weights packed on a stream of their own."""
    from stove_b200 import ops
    for setter, getter in ((ops.set_dyn_stream, ops.dyn_stream_enabled), (ops.set_scene_seq, ops.scene_seq_enabled)):
        assert getter() is True
        assert setter(False) is True and getter() is False
        assert setter(True) is False and getter() is True

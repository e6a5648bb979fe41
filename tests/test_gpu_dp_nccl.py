"""NCCL data parallelism on real GPUs (skipped with fewer than two): sharded Stove gradients == the single-rank
gradient of the concatenated batch, through the overlapped piecewise exchange of stove_b200.dp."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_sharded_gradient_equals_single_rank_gradient():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29671',
                          os.path.join(root, 'tests', 'dp_nccl_worker.py')],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith('DPCHECK ')][-1]
    d = json.loads(line[len('DPCHECK '):])
    assert d['overlap_active'] and d['graph_overlap_active']
    assert d['max_rel_err'] < d['tolerance'], d
    assert d['overlapped_vs_single_allreduce'] < 1e-5, d
    assert d['graph_replay_finite']

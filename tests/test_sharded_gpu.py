"""GPU tests of the multi-GPU front door `batch.causal_impact_sharded` (contiguous blocks of series per
rank, all chains of a series on one GPU, NO data-path collective, one NCCL all-gather of the per-series summaries).
Reference side: one series per `CausalImpact(...)` call, /root/reference/causalimpact/model.py:361-364."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import synth_problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _data(N=12, T=60, Tf=10, k=3, S=7):
    y, ypost, X = synth_problem(N, T, Tf, k, S, 1, seed=3)
    return y, ypost, X


def test_sharded_single_rank_equals_batch():
    from tfcausalimpact_b200.batch import causal_impact_batch, causal_impact_sharded
    y, ypost, X = _data()
    kw = dict(fit_method='vi', nseasons=7, niter=200, seed=4)
    a = causal_impact_sharded(y, ypost, X, **kw)
    b = causal_impact_batch(y, ypost, X, **kw)
    torch.testing.assert_close(a['summary'], b['summary'], rtol=0, atol=0)
    torch.testing.assert_close(a['p_value'], b['p_value'], rtol=0, atol=0)


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import torch, torch.distributed as dist
from helpers import synth_problem
from tfcausalimpact_b200 import distributed as D
from tfcausalimpact_b200.batch import causal_impact_batch, causal_impact_sharded
y, ypost, X = synth_problem(13, 60, 10, 3, 7, 1, seed=3)          # 13 series: ragged blocks (7 + 6)
kw = dict(fit_method='hmc', nseasons=7, niter=200, chains=2, num_results=20, num_warmup=10, seed=4)
res = causal_impact_sharded(y, ypost, X, **kw)
rank, world = dist.get_rank(), dist.get_world_size()
assert dist.get_backend() == 'nccl' and world == 2
lo, hi = D.shard_range(13, rank, world)
loc = causal_impact_batch(y[lo:hi], ypost[lo:hi], X[lo:hi], device=torch.device('cuda', int(os.environ['LOCAL_RANK'])), **kw)
assert res['summary'].shape == (13, 10, 2) and res['p_value'].shape == (13,)
assert torch.equal(res['summary'][lo:hi], loc['summary']) and torch.equal(res['p_value'][lo:hi], loc['p_value'])
assert torch.isfinite(res['summary']).all()
chk = res['summary'].double().sum().reshape(1).clone()
both = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(both, chk)
assert float(both[0]) == float(both[1]), 'ranks disagree on the gathered summaries'
dist.destroy_process_group()
print('RANK_OK', rank)
'''


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_sharded_two_ranks_nccl(tmp_path):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / 'worker.py'
    script.write_text(WORKER.format(root=ROOT))
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count('RANK_OK') == 2

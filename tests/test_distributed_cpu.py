"""Multi-process host logic (series sharding + summary gather) with gloo, world_size 2, on CPU."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from tfcausalimpact_b200.distributed import shard_range
    for n in (1, 2, 7, 10, 10000, 10001):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a, b), (c, d) in zip(blocks[:-1], blocks[1:]):
                assert b == c and b - a >= d - c >= b - a - 1


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from tfcausalimpact_b200 import distributed as D
    r, w, _ = D.init_from_env('gloo')
    lo, hi = D.shard_range(n, r, w)
    full = torch.arange(n * 21, dtype=torch.float32).reshape(n, 21)
    got = D.gather_series(full[lo:hi].clone(), n)
    tmax = D.max_over_ranks(float(rank + 1))
    q.put((rank, bool(torch.equal(got, full)), tmax))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize('n', [7, 10])
def test_gather_series_world2_gloo(n):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(t == 2.0 for _, _, t in res)

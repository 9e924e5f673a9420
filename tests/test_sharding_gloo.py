"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: shard
ranges, ragged all-gather of per-scenario results, and the sharded driver with
a stand-in solver object (the CUDA solve itself is covered by the gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gigpower_b200.sharding import shard_range, gather_rows, solve_sharded


def test_shard_ranges_partition_the_batch():
    for S in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(S, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == S
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


class _FakeResult:
    def __init__(self, lo, hi, n_vrows):
        s = torch.arange(lo, hi, dtype=torch.float64)
        self.iterations = (s % 7 + 3).to(torch.int32)
        self.status = (s % 2).to(torch.int32)
        self.Vmag_rows = torch.arange(n_vrows, dtype=torch.float64)[:, None] * 1000 + s[None, :]
        self.loss = torch.complex(s, -s)


class _FakeSolution:
    """Stands in for SolutionFBS on CPU: results are a known function of the
    GLOBAL scenario index carried in the first spu row."""
    n_vrows = 5

    def solve_batch(self, spu_rows, outputs, return_device):
        lo = int(round(spu_rows[0, 0].real)) if spu_rows.shape[1] else 0
        return _FakeResult(lo, lo + spu_rows.shape[1], self.n_vrows)


def _worker(rank, world, port, S, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = shard_range(S, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float64)[None, :] * torch.tensor([[1.0], [2.0]])
        full = gather_rows(local, S)
        ok = torch.equal(full, torch.arange(S, dtype=torch.float64)[None, :] * torch.tensor([[1.0], [2.0]]))
        cplx = gather_rows(torch.complex(local[0], -local[0]), S)
        ok &= torch.equal(cplx, torch.complex(torch.arange(S, dtype=torch.float64), -torch.arange(S, dtype=torch.float64)))
        spu_rows = np.arange(S, dtype=float)[None, :] + 0j
        out = solve_sharded(_FakeSolution(), spu_rows)
        s = torch.arange(S, dtype=torch.float64)
        ok &= torch.equal(out['iterations'], (s % 7 + 3).to(torch.int32))
        ok &= torch.equal(out['status'], (s % 2).to(torch.int32))
        ok &= torch.equal(out['Vmag'], torch.arange(5, dtype=torch.float64)[:, None] * 1000 + s[None, :])
        ok &= torch.equal(out['loss'], torch.complex(s, -s))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,S', [(2, 10), (2, 11), (3, 8), (2, 1)])
def test_gloo_gather_ragged(world, S):
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=5) for _ in range(world))
    assert results == {r: True for r in range(world)}


def _peer_fail_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gigpower_b200 import _cabi
        from gigpower_b200.sharding import PeerGather
        try:
            PeerGather(5, 64, slots=1, device=0)          # no CUDA device here: the allocation fails on every rank
            q.put((rank, 'no error'))
        except _cabi.GpError as e:
            q.put((rank, 'GpError' if 'rank 0' in str(e) and 'rank 1' in str(e) else str(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a box without a GPU')
def test_peer_gather_setup_fails_collectively_without_cuda():
    """PeerGather's set-up errors are collective (all ranks raise, nobody is left waiting in an exchange):
    the product path has no CPU fallback, so without a CUDA device every rank gets a GpError."""
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_fail_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(q.get(timeout=5) for _ in range(2)) == {0: 'GpError', 1: 'GpError'}

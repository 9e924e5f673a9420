"""GPU tests of the fused all-gather: the solve kernels store |V| and losses straight
into every rank's gather buffer over peer-mapped memory (GpSolveExtras.n_peers / gigpower_b200.sharding.
PeerGather), or the copy engines push finished rows.

One GPU is enough: with world = 1 the only destination is the rank's own buffer, and the two-rank test
runs both ranks on cuda:0 (CUDA IPC maps an allocation of another process on the same device exactly as
it maps one of a peer device; the control plane is gloo because NCCL refuses two ranks on one GPU)."""
import os
import socket

import numpy as np
import pytest

from helpers import ZIPS, feeder_path, rel_err

pytestmark = pytest.mark.gpu
F13 = 'IEEE_13_Bus_allwye_noxfm_noreg'


def _mult(S, n, seed=7):
    return np.random.Generator(np.random.PCG64(seed)).uniform(0.5, 1.5, size=(S, n))


@pytest.mark.parametrize('kernel,S', [(5, 8192 + 77), (3, 3000)])
def test_single_rank_gather_equals_the_local_outputs(kernel, S):
    import torch
    from gigpower_b200 import _cabi
    from gigpower_b200.sharding import PeerGather
    from gigpower_b200.solution_fbs import SolutionFBS
    s = SolutionFBS(feeder_path(F13), zip_v=np.asarray(ZIPS['test_zip']))
    s._dev()
    if kernel == 5:
        assert s.specialise()
    _cabi.check(_cabi.gp_set_option(b'fbs_kernel', kernel))
    pg = PeerGather(s.tables.n_vrows, S, slots=2)
    try:
        mult = _mult(S, s.circuit.loads.num_elements)
        r = s.solve_batch(load_mult=mult, outputs=('Vmag', 'loss'), return_device=True, gather=(pg, 1))
        pg.wait()
        assert torch.equal(pg.vmag_all(1)[0], r.Vmag_rows)
        assert torch.equal(pg.loss_all(1)[0], r.loss)
        # gather only (no local |V| / loss buffers at all): same values
        pg.vmag_all(0).zero_()
        r0 = s.solve_batch(load_mult=mult, outputs=(), return_device=True, gather=(pg, 0))
        pg.wait()
        assert r0.Vmag_rows is None and torch.equal(pg.vmag_all(0)[0], r.Vmag_rows)
        assert torch.equal(pg.loss_all(0)[0], r.loss)
        # copy-engine transport
        pg.vmag_all(0).zero_()
        pg.loss_all(0).zero_()
        pg.push(0, r.Vmag_rows, r.loss)
        pg.wait()
        assert torch.equal(pg.vmag_all(0)[0], r.Vmag_rows) and torch.equal(pg.loss_all(0)[0], r.loss)
        with pytest.raises(ValueError):
            s.solve_batch(load_mult=mult[:10], gather=(pg, 0))
    finally:
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 0))
        pg.close()


def _rank(rank, world, port, S, kernel, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gigpower_b200 import _cabi
        from gigpower_b200.sharding import PeerGather, shard_range
        from gigpower_b200.solution_fbs import SolutionFBS
        torch.cuda.set_device(0)
        s = SolutionFBS(feeder_path(F13), zip_v=np.asarray(ZIPS['test_zip']), device=0)
        s._dev()
        if kernel == 5:
            assert s.specialise()
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', kernel))
        mult = _mult(S, s.circuit.loads.num_elements)
        lo, hi = shard_range(S, rank, world)
        n = hi - lo
        pg = PeerGather(s.tables.n_vrows, n, slots=1)
        s.solve_batch(load_mult=mult[lo:hi], outputs=(), return_device=True, gather=(pg, 0))
        pg.wait()
        got_v = torch.cat([pg.vmag_all(0)[r] for r in range(world)], dim=1).cpu().numpy()
        got_l = torch.cat([pg.loss_all(0)[r] for r in range(world)], dim=0).cpu().numpy()
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
        full = s.solve_batch(load_mult=mult, outputs=('Vmag', 'loss'))
        ok = rel_err(got_v, full.Vmag_rows) <= 1e-12 and rel_err(got_l, full.loss) <= 1e-12
        # second transport: DMA push of this rank's rows into a zeroed slot
        pg.wait()
        pg.vmag_all(0).zero_()
        pg.wait()
        mine = s.solve_batch(load_mult=mult[lo:hi], outputs=('Vmag', 'loss'), return_device=True)
        pg.push(0, mine.Vmag_rows, mine.loss)
        pg.wait()
        got_v2 = torch.cat([pg.vmag_all(0)[r] for r in range(world)], dim=1).cpu().numpy()
        ok = ok and rel_err(got_v2, full.Vmag_rows) <= 1e-12
        # third: the solve writes this rank's block of its OWN buffer, one DMA per peer distributes it
        pg.wait()
        pg.vmag_all(0).zero_()
        pg.loss_all(0).zero_()
        pg.wait()
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', kernel))
        s.solve_batch(load_mult=mult[lo:hi], outputs=(), return_device=True, gather=(pg, 0, 'own'))
        pg.push_own(0)
        pg.wait()
        got_v3 = torch.cat([pg.vmag_all(0)[r] for r in range(world)], dim=1).cpu().numpy()
        got_l3 = torch.cat([pg.loss_all(0)[r] for r in range(world)], dim=0).cpu().numpy()
        ok = ok and rel_err(got_v3, full.Vmag_rows) <= 1e-12 and rel_err(got_l3, full.loss) <= 1e-12
        pg.close()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('kernel,S', [(5, 2 * 8192), (3, 2 * 1500)])
def test_two_ranks_gather_through_peer_memory(kernel, S):
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(('127.0.0.1', 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(r, 2, port, S, kernel, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert dict(q.get(timeout=5) for _ in range(2)) == {0: True, 1: True}

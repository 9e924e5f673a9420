"""GPU parity tests of the batched FBS path (through the C-ABI) against the
oracle and the golden vectors.  Bar (BASELINE.json north_star): complex bus
voltages within 1e-9 relative, identical iteration counts."""
import numpy as np
import pytest

from helpers import (ZIPS, ALL_FEEDERS, LINES_ONLY, feeder_path, load_model, snapshots, scenarios, rel_err,
                     recorded, pinned_34, pinned_34_taps, assert_close)

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _fbs(feeder, zname='test_zip', **kw):
    from gigpower_b200.solution_fbs import SolutionFBS
    return SolutionFBS(feeder_path(feeder), zip_v=np.asarray(ZIPS[zname]), **kw)


def _oracle(feeder, zname='test_zip'):
    from oracle.gp_oracle import FBSOracle
    return FBSOracle(load_model(feeder), ZIPS[zname])


@pytest.mark.parametrize('zname', list(ZIPS))
@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_snapshot_matches_reference(feeder, zname):
    snap = snapshots()
    s = _fbs(feeder, zname)
    s.solve()
    k = f'{feeder}/{zname}/FBS/'
    assert s.iterations == int(snap[k + 'iterations'])
    assert s.converged == bool(snap[k + 'converged'])
    assert abs(s.convergence_diff - float(snap[k + 'convergence_diff'])) <= 1e-12
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx'):
        assert getattr(s, p).shape == snap[k + p].shape, p
        assert_close(getattr(s, p), snap[k + p], TOL, p)
    # accessors: reference shapes (tests/test_api.py:47-71)
    for p in s.SOLUTION_PARAMS:
        rows, cols = s.get_data_frame(p, orient='rows'), s.get_data_frame(p, orient='cols')
        assert not rows.empty and rows.shape[1] == 3 and cols.shape[0] == 3
        assert rows.shape == cols.shape[::-1]
    assert list(s.get_V().index) == s.circuit.buses.all_names()
    assert s.get_I().shape[0] == s.circuit.lines.num_elements
    # a second solve() is a no-op, as in the reference (iterations / Vtest persist)
    V1 = s.V.copy()
    s.solve()
    assert np.array_equal(V1, s.V) and s.iterations == int(snap[k + 'iterations'])


@pytest.mark.parametrize('zname', list(ZIPS))
def test_34_bus_snapshot_at_the_taps_of_the_recorded_tables(zname):
    """IEEE_34_Bus_allwye.dss with the six taps recovered by oracle/pin_taps_34.py: the
    unmodified reference's full-precision results, and gigpower's own recorded tables."""
    z, rec = pinned_34(), recorded()
    s = _fbs('IEEE_34_Bus_allwye', zname, taps=pinned_34_taps())
    s.solve()
    k = f'{zname}/FBS/'
    assert s.iterations == int(z[k + 'iterations'])
    assert abs(s.convergence_diff - float(z[k + 'convergence_diff'])) <= 1e-12
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx'):
        assert getattr(s, p).shape == z[k + p].shape, p
        assert rel_err(getattr(s, p), z[k + p]) <= TOL, p
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx'):
        table = rec[f'IEEE_34_Bus_allwye/{zname}/FBS/{p}']
        frame = s.get_data_frame(p, orient='rows')
        assert list(table) == list(frame.index)
        want = np.array([[complex(*x) for x in table[n]] for n in table])
        got = frame.to_numpy()
        assert np.abs(got.real - want.real).max() <= 0.005 + 1e-9, p
        assert np.abs(got.imag - want.imag).max() <= 0.005 + 1e-9, p


@pytest.mark.parametrize('feeder', LINES_ONLY)
def test_golden_scenarios(feeder):
    sc = scenarios()
    s = _fbs(feeder)
    r = s.solve_batch(load_mult=sc[feeder + '/mult'], outputs=('V', 'I'))
    assert list(r.iterations) == list(sc[feeder + '/FBS/iterations'])
    assert rel_err(r.V, sc[feeder + '/FBS/V']) <= TOL
    assert rel_err(r.I, sc[feeder + '/FBS/I']) <= TOL


@pytest.mark.parametrize('feeder,S,seed', [('IEEE_13_Bus_allwye_noxfm_noreg', 4096, 1),
                                            ('IEEE_34_Bus_allwye_noxfm_noreg', 1000, 2),
                                            ('IEEE_37_Bus_allwye_noxfm_noreg', 1027, 3),
                                            ('IEEE_13_Bus_allwye', 300, 4),
                                            ('IEEE_37_Bus_allwye', 129, 5)])
def test_random_batch_matches_oracle(feeder, S, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    o = _oracle(feeder)
    mult = rng.uniform(0.5, 1.5, size=(S, len(o.c.load_names)))
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    want = o.solve(spu)
    s = _fbs(feeder)
    r = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag', 'sV', 'Stx', 'Srx', 'loss'))
    assert np.array_equal(r.iterations, want['iterations'])
    assert np.array_equal(r.converged, want['converged'])
    for p in ('V', 'I', 'Vmag', 'sV', 'Stx', 'Srx'):
        assert_close(getattr(r, p), want[p], TOL, p)
    loss = (want['Stx'] - want['Srx']).sum(axis=(1, 2))
    assert rel_err(r.losses, loss) <= TOL
    assert np.abs(r.convergence_diff - want['convergence_diff']).max() <= 1e-12
    # the same scenarios given as the reference's per-bus spu matrix
    r2 = s.solve_batch(spu=spu, outputs=('V',))
    assert np.array_equal(r2.iterations, r.iterations) and rel_err(r2.V, r.V) <= 1e-13


def test_full_size_config_c2_matches_oracle():
    """BASELINE config C2: 13-bus, 65,536 scenarios, m ~ U(0.5, 1.5), seed 13."""
    feeder, S = 'IEEE_13_Bus_allwye_noxfm_noreg', 65536
    rng = np.random.Generator(np.random.PCG64(13))
    o = _oracle(feeder)
    mult = rng.uniform(0.5, 1.5, size=(S, len(o.c.load_names)))
    want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    r = _fbs(feeder).solve_batch(load_mult=mult, outputs=('V', 'Vmag', 'loss'))
    assert np.array_equal(r.iterations, want['iterations'])
    assert r.converged.all()
    assert rel_err(r.V, want['V']) <= TOL
    assert rel_err(r.Vmag, want['Vmag']) <= TOL
    # size-independent property: losses are non-negative real power, and KCL at the root:
    assert (r.losses.real > 0).all()


def test_edge_cases_empty_ragged_single():
    feeder = LINES_ONLY[0]
    s = _fbs(feeder)
    o = _oracle(feeder)
    nload = len(o.c.load_names)
    r0 = s.solve_batch(load_mult=np.zeros((0, nload)))
    assert r0.S == 0 and r0.V.shape[0] == 0
    for S in (1, 2, 127, 130):
        mult = np.linspace(0.5, 1.5, S * nload).reshape(S, nload)
        want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
        r = s.solve_batch(load_mult=mult)
        assert np.array_equal(r.iterations, want['iterations'])
        assert rel_err(r.V, want['V']) <= TOL
    # zero load: converges in 2 sweeps to the flat voltage profile
    r = s.solve_batch(load_mult=np.zeros((3, nload)))
    want = o.solve(o.c.spu_matrix(o.c.load_kw * 0, o.c.load_kvar * 0)[None].repeat(3, 0))
    assert np.array_equal(r.iterations, want['iterations'])
    assert rel_err(r.V, want['V']) <= TOL


def test_non_convergent_scenario_reports_status_not_error():
    """34-bus at 2.0x load does not converge in the reference:
    flagged per scenario, neighbours unaffected."""
    feeder = 'IEEE_34_Bus_allwye_noxfm_noreg'
    o = _oracle(feeder)
    nload = len(o.c.load_names)
    mult = np.ones((4, nload))
    mult[1] = 2.0
    mult[2] = 1.5      # 59 iterations in the reference
    want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    r = _fbs(feeder).solve_batch(load_mult=mult)
    assert np.array_equal(r.iterations, want['iterations'])
    assert r.iterations[1] == 100 and r.status[1] in (1, 2) and not r.converged[1]
    assert r.converged[[0, 2, 3]].all()
    assert rel_err(r.V[[0, 2, 3]], want['V'][[0, 2, 3]]) <= TOL


def test_host_buffer_entry_point_matches_device_path():
    import torch
    feeder, S = LINES_ONLY[0], 20000
    s = _fbs(feeder)
    rng = np.random.Generator(np.random.PCG64(7))
    rows = s.load_weights() @ rng.uniform(0.5, 1.5, size=(S, s.circuit.loads.num_elements)).T
    r = s.solve_batch(spu_rows=rows, outputs=('V', 'I', 'Vmag', 'loss'))
    spu_h = torch.from_numpy(np.ascontiguousarray(rows)).pin_memory()
    t = s.tables
    V = torch.empty((t.n_vrows, S), dtype=torch.complex128).pin_memory()
    I = torch.empty((t.n_irows, S), dtype=torch.complex128).pin_memory()
    Vmag = torch.empty((t.n_vrows, S), dtype=torch.float64).pin_memory()
    loss = torch.empty(S, dtype=torch.complex128).pin_memory()
    iters, status, diff = s.solve_batch_host(spu_h, V, I, Vmag, loss)
    assert np.array_equal(iters, r.iterations) and np.array_equal(status, r.status)
    assert np.array_equal(V.numpy(), r.V_rows) and np.array_equal(I.numpy(), r.I_rows)
    assert np.array_equal(Vmag.numpy(), r.Vmag_rows) and np.array_equal(loss.numpy(), r.loss)


def test_errors_are_python_exceptions():
    from gigpower_b200._cabi import GpError
    s = _fbs(LINES_ONLY[0])
    with pytest.raises(ValueError):
        s.solve_batch(spu_rows=np.zeros((3, 4), dtype=complex))
    s._dev()
    from gigpower_b200 import _cabi
    with pytest.raises(GpError):
        _cabi.check(_cabi.gp_fbs_solve_batch(s._feeder.handle, 4, 2, None, None, None, None, None, None,
                                             1e-6, 100, None))


KERNELS = {0: 'automatic', 1: 'fused depth-first', 2: 'phase-parallel', 3: 'two-sweep streaming', 4: 'on-chip',
           5: 'feeder-specialised on-chip', 6: 'tree walk'}


@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye', 'IEEE_13_Bus_allwye_noxfm_noreg',
                                    'IEEE_34_Bus_allwye_noxfm_noreg', 'IEEE_34_Bus_allwye', 'IEEE_37_Bus_allwye'])
def test_all_fbs_kernels_agree(feeder):
    """The library has four FBS kernels (on-chip and two-sweep streaming = the automatic choices,
    fused depth-first walk and phase-parallel lanes = options); each must meet the oracle, and
    they must agree with each other to rounding."""
    from gigpower_b200 import _cabi
    rng = np.random.Generator(np.random.PCG64(21))
    o = _oracle(feeder)
    mult = rng.uniform(0.5, 1.5, size=(777, len(o.c.load_names)))
    want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    s = _fbs(feeder)
    assert s.specialise(), 'NVRTC / driver module load failed on the GPU box'
    res = {}
    try:
        for mode in KERNELS:
            _cabi.check(_cabi.gp_set_option(b'fbs_kernel', mode))
            if mode == 6 and _cabi.gp_fbs_active_kernel(s._feeder.handle) != 6:
                continue        # a regulator that leaves phases of its bus unregulated: not walked
            res[mode] = s.solve_batch(load_mult=mult, outputs=('V', 'I'))
        # the streaming kernel with every I row travelling (no shared rows)
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
        _cabi.check(_cabi.gp_set_option(b'fbs_share_rows', 0))
        plain = s.solve_batch(load_mult=mult, outputs=('V', 'I'))
    finally:
        _cabi.gp_set_option(b'fbs_kernel', 0)
        _cabi.gp_set_option(b'fbs_share_rows', 1)
    for mode in res:
        assert np.array_equal(res[mode].iterations, want['iterations']), KERNELS[mode]
        assert rel_err(res[mode].V, want['V']) <= TOL and rel_err(res[mode].I, want['I']) <= TOL, KERNELS[mode]
        assert rel_err(res[3].V, res[mode].V) <= 1e-12, KERNELS[mode]
    # the on-chip and the streaming kernel run the same arithmetic in the same order: bit-identical
    # (with shared I rows the streaming kernel takes the generic body for steps that lose a row, which
    # adds the load current and the children's currents in the other order: equal to rounding)
    assert np.array_equal(plain.V, res[4].V) and np.array_equal(plain.I, res[4].I)
    assert np.array_equal(plain.iterations, res[3].iterations) and np.array_equal(plain.status, res[3].status)
    assert rel_err(res[3].V, plain.V) <= 1e-13 and rel_err(res[3].I, plain.I) <= 1e-13


def test_onchip_kernel_full_batch_and_relaunch():
    """On-chip kernel at BASELINE C2 size: persistent warps pull 32-scenario tiles from a device
    counter that the last warp re-arms, so back-to-back launches must give identical results;
    ragged last tile; results bit-identical to the streaming kernel."""
    from gigpower_b200 import _cabi
    feeder, S = 'IEEE_13_Bus_allwye_noxfm_noreg', 65536 + 17
    rng = np.random.Generator(np.random.PCG64(5))
    s = _fbs(feeder)
    mult = rng.uniform(0.5, 1.5, size=(S, s.circuit.loads.num_elements))
    s._dev()
    assert _cabi.gp_fbs_onchip_warps(s._feeder.handle) >= 4
    assert _cabi.gp_fbs_active_kernel(s._feeder.handle) in (3, 6)      # nothing specialised yet: a streaming kernel
    try:
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
        shared = s.solve_batch(load_mult=mult, outputs=('V', 'I'))
        _cabi.check(_cabi.gp_set_option(b'fbs_share_rows', 0))         # bit-identity holds row for row
        ref = s.solve_batch(load_mult=mult, outputs=('V', 'I'))
        _cabi.check(_cabi.gp_set_option(b'fbs_share_rows', 1))
        assert np.array_equal(shared.iterations, ref.iterations) and np.array_equal(shared.status, ref.status)
        assert rel_err(shared.V, ref.V) <= 1e-13 and rel_err(shared.I, ref.I) <= 1e-13
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 4))
        for _ in range(3):
            r = s.solve_batch(load_mult=mult, outputs=('V', 'I'))
            assert np.array_equal(r.iterations, ref.iterations) and np.array_equal(r.status, ref.status)
            assert np.array_equal(r.V, ref.V) and np.array_equal(r.I, ref.I)
            assert np.array_equal(r.convergence_diff, ref.convergence_diff)
        for S2 in (1, 31, 33, 4097):
            r = s.solve_batch(load_mult=mult[:S2], outputs=('V',))
            assert np.array_equal(r.V, ref.V[:S2]) and np.array_equal(r.iterations, ref.iterations[:S2])
        # the batch above was large enough to compile the feeder-specialised kernel: automatic = 5 now
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 0))
        assert s._specialised and _cabi.gp_fbs_active_kernel(s._feeder.handle) == 5
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
        ref2 = s.solve_batch(load_mult=mult, outputs=('Vmag', 'loss'))
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 0))
        for _ in range(2):
            r = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag', 'loss'))   # |V|, loss fused in the kernel
            assert np.array_equal(r.iterations, ref.iterations) and np.array_equal(r.status, ref.status)
            assert rel_err(r.V, ref.V) <= 1e-13 and rel_err(r.I, ref.I) <= 1e-13
            assert rel_err(r.Vmag, ref2.Vmag) <= 1e-13 and rel_err(r.losses, ref2.losses) <= 1e-13
        for S2 in (1, 31, 33, 4097):
            r = s.solve_batch(load_mult=mult[:S2], outputs=('V',))
            assert rel_err(r.V, ref.V[:S2]) <= 1e-13 and np.array_equal(r.iterations, ref.iterations[:S2])
    finally:
        _cabi.gp_set_option(b'fbs_kernel', 0)
        _cabi.gp_set_option(b'fbs_share_rows', 1)


def test_specialised_kernel_restarts_diverging_scenarios_exactly():
    """The specialised kernel runs branch-free arithmetic and restarts a scenario with the exact
    nan_to_num / division semantics when a current goes non-finite: heavily overloaded scenarios
    (which blow up) must report the same iterations and status as the table-driven kernel, and
    their well-behaved neighbours the same voltages."""
    from gigpower_b200 import _cabi
    feeder = 'IEEE_13_Bus_allwye_noxfm_noreg'
    s = _fbs(feeder)
    assert s.specialise()
    nload = s.circuit.loads.num_elements
    rng = np.random.Generator(np.random.PCG64(99))
    mult = rng.uniform(0.5, 1.5, size=(512, nload))
    mult[5] *= 40.0
    mult[100] *= 1e3
    mult[257] *= 1e6
    mult[300] = 0.0
    try:
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
        ref = s.solve_batch(load_mult=mult, outputs=('V', 'I'))
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 5))
        r = s.solve_batch(load_mult=mult, outputs=('V', 'I'))
    finally:
        _cabi.gp_set_option(b'fbs_kernel', 0)
    assert (ref.status != 0).sum() >= 2, 'the overloaded scenarios were expected to fail'
    assert np.array_equal(r.iterations, ref.iterations) and np.array_equal(r.status, ref.status)
    ok = ref.status == 0
    assert rel_err(r.V[ok], ref.V[ok]) <= 1e-12 and rel_err(r.I[ok], ref.I[ok]) <= 1e-12
    bad = ~ok
    fin = np.isfinite(ref.V[bad]) & np.isfinite(r.V[bad])
    assert np.array_equal(np.isfinite(ref.V[bad]), np.isfinite(r.V[bad]))
    assert rel_err(r.V[bad][fin], ref.V[bad][fin]) <= 1e-9


def test_host_entry_point_one_launch_path_matches_staged_path():
    """gp_fbs_solve_batch_host with the specialised kernel and pinned buffers runs as ONE launch that
    reads spu and writes the results over PCIe itself; it must give exactly what the staged
    (copy - solve - copy) path gives, with any subset of the optional outputs, and pageable buffers
    must fall back to the staged path."""
    import torch
    from gigpower_b200 import _cabi
    feeder, S = 'IEEE_13_Bus_allwye_noxfm_noreg', 30000 + 7
    s = _fbs(feeder)
    assert s.specialise()
    rng = np.random.Generator(np.random.PCG64(17))
    rows = np.ascontiguousarray(s.load_weights() @ rng.uniform(0.5, 1.5, size=(S, s.circuit.loads.num_elements)).T)
    t = s.tables
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    spu_h = torch.from_numpy(rows).pin_memory()
    out = {}
    try:
        for zc in (0, 1, 2):       # staged; one launch with the read-ahead build of the kernel; one launch without
            _cabi.check(_cabi.gp_set_option(b'host_zero_copy', min(zc, 1)))
            _cabi.check(_cabi.gp_set_option(b'host_read_ahead', 0 if zc == 2 else 1))
            V, I = pin((t.n_vrows, S), torch.complex128), pin((t.n_irows, S), torch.complex128)
            Vmag, loss = pin((t.n_vrows, S), torch.float64), pin(S, torch.complex128)
            n0 = _cabi.gp_launch_count()
            it, st, df = s.solve_batch_host(spu_h, V, I, Vmag, loss)
            out[zc] = (it.copy(), st.copy(), df.copy(), V.numpy().copy(), I.numpy().copy(), Vmag.numpy().copy(),
                       loss.numpy().copy(), _cabi.gp_launch_count() - n0)
    finally:
        _cabi.gp_set_option(b'host_zero_copy', 1)
        _cabi.gp_set_option(b'host_read_ahead', 1)
    assert out[1][7] == 1 and out[2][7] == 1 and out[0][7] >= 1
    for a, b, c in zip(out[0][:7], out[1][:7], out[2][:7]):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    # only V requested; and pageable (not pinned) buffers -> staged path, same numbers
    V2 = pin((t.n_vrows, S), torch.complex128)
    it2, st2, _ = s.solve_batch_host(spu_h, V2)
    assert np.array_equal(V2.numpy(), out[1][3]) and np.array_equal(it2, out[1][0])
    V3 = np.empty((t.n_vrows, S), dtype=complex)
    it3, _, _ = s.solve_batch_host(rows, V3)
    assert np.array_equal(V3, out[1][3]) and np.array_equal(it3, out[1][0])


@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_13_Bus_allwye', 'IEEE_34_Bus_allwye_noxfm_noreg',
                                    'IEEE_37_Bus_allwye'])
def test_specialised_kernel_with_i_rows_in_shared_or_global_memory(feeder):
    """The specialised kernel keeps the I rows either in shared memory or in the (L2-resident) I
    array ("jit_i_global"); both placements must reproduce the streaming kernel, through the device
    entry point and through the one-launch host entry point, with and without an I output."""
    import torch
    from gigpower_b200 import _cabi
    rng = np.random.Generator(np.random.PCG64(31))
    ref = None
    try:
        for ig in (0, 1):
            _cabi.check(_cabi.gp_set_option(b'jit_i_global', ig))
            s = _fbs(feeder)
            mult = np.random.Generator(np.random.PCG64(31)).uniform(0.5, 1.5, size=(2500, s.circuit.loads.num_elements))
            assert s.specialise()
            if ref is None:
                _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
                ref = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag', 'loss'))
            _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 5))
            for _ in range(2):
                r = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag', 'loss'))
                assert np.array_equal(r.iterations, ref.iterations) and np.array_equal(r.status, ref.status)
                assert rel_err(r.V, ref.V) <= 1e-12 and rel_err(r.I, ref.I) <= 1e-12
                assert rel_err(r.Vmag, ref.Vmag) <= 1e-12 and rel_err(r.losses, ref.losses) <= 1e-12
            rows = np.ascontiguousarray(s.load_weights() @ mult.T)
            t = s.tables
            spu_h = torch.from_numpy(rows).pin_memory()
            V = torch.empty((t.n_vrows, 2500), dtype=torch.complex128).pin_memory()
            I = torch.empty((t.n_irows, 2500), dtype=torch.complex128).pin_memory()
            it, st, df = s.solve_batch_host(spu_h, V, I)
            assert np.array_equal(it, ref.iterations)
            assert np.array_equal(V.numpy(), np.asarray(r.V_rows)) and np.array_equal(I.numpy(), np.asarray(r.I_rows))
            V2 = torch.empty((t.n_vrows, 2500), dtype=torch.complex128).pin_memory()
            s.solve_batch_host(spu_h, V2)
            assert np.array_equal(V2.numpy(), V.numpy())
    finally:
        _cabi.gp_set_option(b'jit_i_global', -1)
        _cabi.gp_set_option(b'fbs_kernel', 0)


@pytest.mark.parametrize('kernel', [5, 4])
def test_chunked_host_path_runs_persistent_kernels_on_several_streams(kernel):
    """The staged host path cuts a batch into chunks on up to four streams.  The persistent kernels
    (feeder-specialised, table-driven on-chip) pull tiles from scheduler words that are only right
    for ONE kernel in flight, so each stream has its own pair: four chunks in flight must give
    exactly the one-chunk result (every tile solved, once)."""
    import torch
    from gigpower_b200 import _cabi
    feeder, S = 'IEEE_13_Bus_allwye_noxfm_noreg', 4 * 9000 + 77
    s = _fbs(feeder)
    assert s.specialise()
    rng = np.random.Generator(np.random.PCG64(71))
    rows = np.ascontiguousarray(s.load_weights() @ rng.uniform(0.5, 1.5, size=(S, s.circuit.loads.num_elements)).T)
    t = s.tables
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    spu_h = torch.from_numpy(rows).pin_memory()
    out = {}
    try:
        _cabi.check(_cabi.gp_set_option(b'host_zero_copy', 0))
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', kernel))
        for chunks in (1, 4, 7):
            _cabi.check(_cabi.gp_set_option(b'host_chunks', chunks))
            for rep in range(3):            # relaunches: the words must have re-armed themselves
                V, I = pin((t.n_vrows, S), torch.complex128), pin((t.n_irows, S), torch.complex128)
                V.fill_(float('nan')), I.fill_(float('nan'))
                Vmag, loss = pin((t.n_vrows, S), torch.float64), pin(S, torch.complex128)
                it, st, df = s.solve_batch_host(spu_h, V, I, Vmag, loss)
                res = (it.copy(), st.copy(), df.copy(), V.numpy().copy(), I.numpy().copy(), Vmag.numpy().copy(),
                       loss.numpy().copy())
                if chunks == 1 and rep == 0:
                    out = res
                    assert np.isfinite(res[3]).all() and (res[1] == 0).all()
                else:
                    for a, b in zip(out, res):
                        assert np.array_equal(a, b), f'{chunks} chunks, repeat {rep}'
    finally:
        _cabi.gp_set_option(b'host_zero_copy', 1)
        _cabi.gp_set_option(b'host_chunks', 0)
        _cabi.gp_set_option(b'fbs_kernel', 0)


@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_13_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg'])
def test_calc_methods_match_the_reference(feeder):
    """Solution.calc_sV / calc_sV(bus) / calc_Vmag / calc_Stx / calc_Srx (solution.py:177-233) on an object
    whose V, I, spu and wpu were set by the caller: against the unmodified reference's own methods on
    the same arbitrary state (tests/golden/ref_calc.npz, oracle/make_golden_calc.py)."""
    import os
    from helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'ref_calc.npz'))
    s = _fbs(feeder)
    k = feeder + '/'
    s.V, s.I, s.wpu, s.spu = g[k + 'V'].copy(), g[k + 'I'].copy(), g[k + 'wpu'].copy(), g[k + 'spu'].copy()
    s.calc_sV()
    s.calc_Vmag()
    s.calc_Stx()
    s.calc_Srx()
    nl = s.circuit.lines.num_elements
    assert np.abs(s.sV - g[k + 'sV']).max() <= 1e-13 and np.abs(s.Vmag - g[k + 'Vmag']).max() <= 1e-13
    assert np.abs(s.Stx[:nl] - g[k + 'Stx']).max() <= 1e-13 and np.abs(s.Srx[:nl] - g[k + 'Srx']).max() <= 1e-13
    s.V = g[k + 'bus_V'].copy()
    s.calc_sV(s.circuit.buses.get_element(str(g[k + 'bus_name'])))
    assert np.abs(s.sV - g[k + 'bus_sV']).max() <= 1e-13


@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye', 'IEEE_34_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg', 'ieee123'])
def test_depth_first_order_is_bit_identical_to_two_sweeps(feeder):
    """The streaming kernel's depth-first step order (fbs_order = 1) runs the same steps on the same values
    as the two full sweeps (fbs_order = 0): results must be bit-identical, with and without shared I rows,
    and also for warm starts and per-scenario taps (the EX instance)."""
    import os
    from gigpower_b200 import _cabi
    from helpers import REPO
    if feeder == 'ieee123':
        from gigpower_b200.solution_fbs import SolutionFBS
        s = SolutionFBS(os.path.join(REPO, 'gigpower_b200', 'feeders', 'ieee123_allwye_noxfm_noreg.dss'))
    else:
        s = _fbs(feeder)
    rng = np.random.Generator(np.random.PCG64(77))
    mult = rng.uniform(0.4, 1.4, size=(3000, s.circuit.loads.num_elements))
    out = {}
    try:
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
        for order in (0, 1):
            for share in (0, 1):
                _cabi.check(_cabi.gp_set_option(b'fbs_order', order))
                _cabi.check(_cabi.gp_set_option(b'fbs_share_rows', share))
                r = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag', 'loss'))
                out[order, share] = (r.iterations, r.status, np.asarray(r.V_rows), np.asarray(r.I_rows), r.losses)
                w = s.solve_batch(load_mult=mult * 1.05, outputs=('V', 'I'), return_device=True)
                w2 = s.solve_batch(load_mult=mult, outputs=('V', 'I'), warm_start=w)
                out['warm', order, share] = (w2.iterations, np.asarray(w2.V_rows), np.asarray(w2.I_rows))
        for share in (0, 1):
            for a, b in zip(out[0, share], out[1, share]):
                assert np.array_equal(a, b), share
            for a, b in zip(out['warm', 0, share], out['warm', 1, share]):
                assert np.array_equal(a, b), ('warm', share)
    finally:
        _cabi.gp_set_option(b'fbs_kernel', 0)
        _cabi.gp_set_option(b'fbs_order', 1)
        _cabi.gp_set_option(b'fbs_share_rows', 1)


@pytest.mark.parametrize('feeder', ['ieee123', 'synth2500', 'IEEE_37_Bus_allwye_noxfm_noreg', 'IEEE_13_Bus_allwye'])
def test_tree_walk_kernel_matches_oracle_and_two_sweep_kernel(feeder):
    """fbs_walk_kernel in each of its register budgets (fbs_walk_blocks 4/5/6/8) on ragged batch sizes:
    oracle at 1e-9 with equal iteration counts, the two-sweep streaming kernel at 1e-12 (the only
    difference is the order in which a bus's child currents are added), status words equal -
    including scenarios that do not converge."""
    import os
    from gigpower_b200 import _cabi
    from gigpower_b200.dss_reader import read_dss
    from gigpower_b200.solution_fbs import SolutionFBS
    from oracle.gp_oracle import FBSOracle
    from helpers import REPO
    files = {'ieee123': 'ieee123_allwye_noxfm_noreg.dss', 'synth2500': 'synth2500_radial.dss'}
    if feeder in files:
        path = os.path.join(REPO, 'gigpower_b200', 'feeders', files[feeder])
        s, o = SolutionFBS(path, zip_v=np.asarray(ZIPS['test_zip'])), FBSOracle(read_dss(path), ZIPS['test_zip'])
    else:
        s, o = _fbs(feeder), _oracle(feeder)
    S = 150 if feeder == 'synth2500' else 1031
    rng = np.random.Generator(np.random.PCG64(66))
    mult = rng.uniform(0.4, 1.3, size=(S, len(o.c.load_names)))
    mult[5] = 6.0           # overloaded: does not converge / diverges
    want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    try:
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
        ref = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag', 'loss'))
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 6))
        assert _cabi.gp_fbs_active_kernel(s._feeder.handle) == 6
        for blocks in (4, 5, 6, 8):
            _cabi.check(_cabi.gp_set_option(b'fbs_walk_blocks', blocks))
            r = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag', 'loss'))
            ok = ref.status == 0
            assert ok.sum() >= S - 3
            assert np.array_equal(r.status, ref.status), blocks
            assert np.array_equal(r.iterations[ok], ref.iterations[ok]), blocks
            assert np.array_equal(r.iterations[ok], want['iterations'][ok]), blocks
            assert rel_err(r.V[ok], want['V'][ok]) <= TOL and rel_err(r.I[ok], want['I'][ok]) <= TOL, blocks
            assert rel_err(r.V[ok], ref.V[ok]) <= 1e-12 and rel_err(r.I[ok], ref.I[ok]) <= 1e-12, blocks
            assert rel_err(r.Vmag[ok], ref.Vmag[ok]) <= 1e-12 and rel_err(r.losses[ok], ref.losses[ok]) <= 1e-12, blocks
    finally:
        _cabi.gp_set_option(b'fbs_kernel', 0)
        _cabi.gp_set_option(b'fbs_walk_blocks', 4)


@pytest.mark.parametrize('feeder', ['IEEE_34_Bus_allwye_noxfm_noreg', 'IEEE_37_Bus_allwye', 'IEEE_13_Bus_allwye', 'IEEE_34_Bus_allwye'])
def test_depth_first_build_of_the_specialised_kernel(feeder):
    """jit_dfs = 1: the specialised kernel walks the tree depth-first with the V rows in the caller's array and only the
    I slots and a small stack in shared memory (more warps per SM on the 34- / 37-bus class).  Same iteration counts and
    status as the streaming kernel, V / I / |V| / losses within 1e-12, through the device entry point (ragged batch, two
    launches in a row) and the host entry point (one launch of the two-sweep build; staged when zero copy is off); diverging
    scenarios take the exact restart path."""
    import torch
    from gigpower_b200 import _cabi
    S = 2500 + 13
    rng = np.random.Generator(np.random.PCG64(41))
    try:
        _cabi.check(_cabi.gp_set_option(b'jit_dfs', 1))
        s = _fbs(feeder)
        mult = rng.uniform(0.5, 1.4, size=(S, s.circuit.loads.num_elements))
        mult[7] = 9.0       # overloaded
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
        ref = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag', 'loss'))
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 0))
        assert s.specialise() and _cabi.gp_fbs_active_kernel(s._feeder.handle) == 5
        ok = ref.status == 0
        assert ok.sum() >= S - 2
        for _ in range(2):
            r = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag', 'loss'))
            assert np.array_equal(r.iterations, ref.iterations) and np.array_equal(r.status, ref.status)
            assert rel_err(r.V[ok], ref.V[ok]) <= 1e-12 and rel_err(r.I[ok], ref.I[ok]) <= 1e-12
            assert rel_err(r.Vmag[ok], ref.Vmag[ok]) <= 1e-12 and rel_err(r.losses[ok], ref.losses[ok]) <= 1e-12
        rows = np.ascontiguousarray(s.load_weights() @ mult.T)
        t = s.tables
        spu_h = torch.from_numpy(rows).pin_memory()
        V = torch.empty((t.n_vrows, S), dtype=torch.complex128).pin_memory()
        I = torch.empty((t.n_irows, S), dtype=torch.complex128).pin_memory()
        n0 = _cabi.gp_launch_count()
        it, st, df = s.solve_batch_host(spu_h, V, I)
        assert _cabi.gp_launch_count() - n0 == 1        # one launch: the two-sweep build, compiled for the host path
        assert np.array_equal(it, ref.iterations) and np.array_equal(st, ref.status)
        assert rel_err(V.numpy()[:, ok], np.asarray(r.V_rows)[:, ok]) <= 1e-12
        assert rel_err(I.numpy()[:, ok], np.asarray(r.I_rows)[:, ok]) <= 1e-12
        # the staged path (copy - solve with the depth-first build - copy) gives the device path's numbers exactly
        _cabi.check(_cabi.gp_set_option(b'host_zero_copy', 0))
        V2 = torch.empty((t.n_vrows, S), dtype=torch.complex128).pin_memory()
        it2, _, _ = s.solve_batch_host(spu_h, V2)
        assert np.array_equal(it2, ref.iterations) and np.array_equal(V2.numpy()[:, ok], np.asarray(r.V_rows)[:, ok])
    finally:
        _cabi.gp_set_option(b'host_zero_copy', 1)
        _cabi.gp_set_option(b'jit_dfs', -1)
        _cabi.gp_set_option(b'fbs_kernel', 0)

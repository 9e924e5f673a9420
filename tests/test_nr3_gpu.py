"""GPU parity tests of the batched NR3 path against the oracle (dense normal
equations, as the reference) and the golden vectors.  Bar: 1e-9 relative on
complex bus voltages, identical Newton iteration counts."""
import numpy as np
import pytest

from helpers import (ZIPS, ALL_FEEDERS, LINES_ONLY, feeder_path, load_model, snapshots, scenarios, rel_err, assert_close,
                     recorded, pinned_34, pinned_34_taps)

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _nr3(feeder, zname='test_zip', **kw):
    from gigpower_b200.solution_nr3 import SolutionNR3
    return SolutionNR3(feeder_path(feeder), zip_v=np.asarray(ZIPS[zname]), **kw)


@pytest.mark.parametrize('zname', list(ZIPS))
@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_snapshot_matches_reference(feeder, zname):
    snap = snapshots()
    s = _nr3(feeder, zname)
    s.solve()
    k = f'{feeder}/{zname}/NR3/'
    assert s.itercount == int(snap[k + 'iterations'])
    assert s.converged is True and s.iterations == 0          # reference quirks
    assert s.XNR.shape == snap[k + 'XNR'].shape
    assert rel_err(s.XNR, snap[k + 'XNR']) <= TOL
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx', 'i_Node'):
        assert getattr(s, p).shape == snap[k + p].shape, p
        assert_close(getattr(s, p), snap[k + p], TOL, p, small=1e-12)
    for p in s.SOLUTION_PARAMS:
        rows, cols = s.get_data_frame(p, orient='rows'), s.get_data_frame(p, orient='cols')
        assert not rows.empty and rows.shape[1] == 3 and cols.shape[0] == 3
    assert list(s.get_V('rows').index) == s.circuit.buses.all_names()
    assert list(s.get_V().index) == ['A', 'B', 'C']          # NR3 defaults to orient='cols' (:30)


@pytest.mark.parametrize('zname', list(ZIPS))
def test_34_bus_snapshot_at_the_taps_of_the_recorded_tables(zname):
    """IEEE_34_Bus_allwye.dss with the six taps recovered by oracle/pin_taps_34.py: the
    unmodified reference's full-precision results, and gigpower's own recorded tables."""
    z, rec = pinned_34(), recorded()
    s = _nr3('IEEE_34_Bus_allwye', zname, taps=pinned_34_taps())
    s.solve()
    k = f'{zname}/NR3/'
    assert s.itercount == int(z[k + 'iterations'])
    assert rel_err(s.XNR, z[k + 'XNR']) <= TOL
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx', 'i_Node'):
        assert getattr(s, p).shape == z[k + p].shape, p
        assert rel_err(getattr(s, p), z[k + p]) <= TOL, p
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx'):
        table = rec[f'IEEE_34_Bus_allwye/{zname}/NR3/{p}']
        frame = s.get_data_frame(p, orient='rows')
        assert list(table) == list(frame.index)
        want = np.array([[complex(*x) for x in table[n]] for n in table])
        got = frame.to_numpy()
        assert np.abs(got.real - want.real).max() <= 0.005 + 1e-9, p
        assert np.abs(got.imag - want.imag).max() <= 0.005 + 1e-9, p


@pytest.mark.parametrize('feeder', LINES_ONLY)
def test_golden_scenarios(feeder):
    sc = scenarios()
    X = sc[feeder + '/NR3/XNR']
    r = _nr3(feeder).solve_batch(load_mult=sc[feeder + '/mult'][:len(X)], outputs=('X',))
    assert list(r.iterations) == list(sc[feeder + '/NR3/iterations'])
    assert rel_err(r.XNR, X) <= TOL


@pytest.mark.parametrize('feeder,S,seed', [('IEEE_13_Bus_allwye_noxfm_noreg', 200, 1),
                                            ('IEEE_34_Bus_allwye_noxfm_noreg', 48, 2),
                                            ('IEEE_37_Bus_allwye_noxfm_noreg', 48, 3)])
def test_random_batch_matches_dense_oracle(feeder, S, seed):
    from oracle.gp_oracle import NR3Oracle
    rng = np.random.Generator(np.random.PCG64(seed))
    o = NR3Oracle(load_model(feeder), ZIPS['test_zip'])
    mult = rng.uniform(0.5, 1.5, size=(S, len(o.c.load_names)))
    want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    r = _nr3(feeder).solve_batch(load_mult=mult,
                                 outputs=('X', 'V', 'I', 'Vmag', 'sV', 'i_Node', 'Stx', 'Srx', 'loss'))
    assert np.array_equal(r.iterations, want['iterations'])
    assert rel_err(r.XNR, want['XNR']) <= TOL
    for p in ('V', 'I', 'Vmag', 'sV', 'i_Node', 'Stx', 'Srx'):
        assert rel_err(getattr(r, p), want[p]) <= TOL, p
    assert rel_err(r.losses, (want['Stx'] - want['Srx']).sum(axis=(1, 2))) <= TOL


@pytest.mark.parametrize('feeder,S,seed', [('IEEE_13_Bus_allwye_noxfm_noreg', 20000, 11),
                                            ('IEEE_37_Bus_allwye_noxfm_noreg', 16384, 37)])
def test_large_batch_matches_tree_oracle(feeder, S, seed):
    """Config C3 size (37-bus, 16,384 scenarios): checked against the NumPy tree
    elimination, which is itself pinned to the dense oracle on CPU."""
    from oracle.nr3_tree import NR3TreeOracle
    rng = np.random.Generator(np.random.PCG64(seed))
    o = NR3TreeOracle(load_model(feeder), ZIPS['test_zip'])
    mult = rng.uniform(0.5, 1.5, size=(S, len(o.c.load_names)))
    want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    r = _nr3(feeder).solve_batch(load_mult=mult, outputs=('X',))
    assert np.array_equal(r.iterations, want['iterations'])
    assert r.converged.all()
    assert rel_err(r.XNR, want['XNR']) <= TOL


def test_edge_cases_and_zero_load():
    from oracle.gp_oracle import NR3Oracle
    feeder = LINES_ONLY[0]
    o = NR3Oracle(load_model(feeder), ZIPS['test_zip'])
    s = _nr3(feeder)
    nload = len(o.c.load_names)
    assert s.solve_batch(load_mult=np.zeros((0, nload))).S == 0
    for S in (1, 65, 130):
        mult = np.linspace(0.0, 1.6, S * nload).reshape(S, nload)
        want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
        r = s.solve_batch(load_mult=mult, outputs=('X',))
        assert np.array_equal(r.iterations, want['iterations'])
        assert rel_err(r.XNR, want['XNR']) <= TOL


@pytest.mark.parametrize('feeder,S,seed', [('IEEE_13_Bus_allwye', 100, 6), ('IEEE_37_Bus_allwye', 40, 7),
                                            ('IEEE_34_Bus_allwye', 40, 8)])
def test_regulator_feeders_batch_matches_dense_oracle(feeder, S, seed):
    """Transformer and regulator edges in the batched Newton kernel."""
    from oracle.gp_oracle import NR3Oracle
    rng = np.random.Generator(np.random.PCG64(seed))
    o = NR3Oracle(load_model(feeder), ZIPS['test_zip'])
    mult = rng.uniform(0.5, 1.3, size=(S, len(o.c.load_names)))
    want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    r = _nr3(feeder).solve_batch(load_mult=mult, outputs=('X', 'V', 'I', 'Stx'))
    assert np.array_equal(r.iterations, want['iterations'])
    assert rel_err(r.XNR, want['XNR']) <= TOL
    assert rel_err(r.V, want['V']) <= TOL and rel_err(r.I, want['I']) <= TOL and rel_err(r.Stx, want['Stx']) <= TOL


@pytest.mark.parametrize('feeder,S', [('IEEE_13_Bus_allwye_noxfm_noreg', 777), ('IEEE_34_Bus_allwye_noxfm_noreg', 333),
                                      ('IEEE_37_Bus_allwye_noxfm_noreg', 1029)])
def test_both_nr3_kernels_meet_the_oracle_and_each_other(feeder, S):
    """nr3_kernel = 1 (one thread per scenario, 6x6 LU in registers) and 2 (three phase lanes per
    scenario, Gauss-Jordan over warp shuffles) solve the same Newton systems: both within 1e-9 of
    the tree oracle with its iteration counts, and within 1e-11 of each other."""
    from gigpower_b200 import _cabi
    from oracle.nr3_tree import NR3TreeOracle
    rng = np.random.Generator(np.random.PCG64(11))
    o = NR3TreeOracle(load_model(feeder), ZIPS['test_zip'])
    mult = rng.uniform(0.4, 1.6, size=(S, len(o.c.load_names)))
    want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    s = _nr3(feeder)
    res = {}
    try:
        for kern in (1, 2):
            _cabi.check(_cabi.gp_set_option(b'nr3_kernel', kern))
            res[kern] = s.solve_batch(load_mult=mult, outputs=('X', 'V', 'I'))
    finally:
        _cabi.gp_set_option(b'nr3_kernel', 0)
    for kern in (1, 2):
        assert np.array_equal(res[kern].iterations, want['iterations']), kern
        assert (res[kern].status == 0).all()
        assert rel_err(res[kern].XNR, want['XNR']) <= TOL, kern
    assert rel_err(res[1].XNR, res[2].XNR) <= 1e-11


@pytest.mark.parametrize('feeder', ['IEEE_37_Bus_allwye_noxfm_noreg', 'IEEE_13_Bus_allwye_noxfm_noreg', 'ieee123'])
def test_phase_kernel_with_staged_rows_is_bit_identical(feeder):
    """The phase-lane kernel with its cp.async ring (nr3_ring 1: next-step rows staged in shared memory) and with
    direct loads (0, with and without the L2 prefetch) does the same arithmetic on the same values: bit-identical X,
    on ragged batch sizes, incl. a feeder with single-phase and two-phase buses."""
    import os
    from gigpower_b200 import _cabi
    from helpers import REPO
    if feeder == 'ieee123':
        from gigpower_b200.solution_nr3 import SolutionNR3
        s = SolutionNR3(os.path.join(REPO, 'gigpower_b200', 'feeders', 'ieee123_allwye_noxfm_noreg.dss'))
    else:
        s = _nr3(feeder)
    rng = np.random.Generator(np.random.PCG64(91))
    mult = rng.uniform(0.4, 1.3, size=(1237, s.circuit.loads.num_elements))
    out = {}
    try:
        _cabi.check(_cabi.gp_set_option(b'nr3_kernel', 2))
        for ring, pf in ((0, 0), (0, 2), (1, 0)):
            _cabi.check(_cabi.gp_set_option(b'nr3_ring', ring))
            _cabi.check(_cabi.gp_set_option(b'nr3_prefetch', pf))
            r = s.solve_batch(load_mult=mult, outputs=('X',))
            out[ring, pf] = (r.iterations, r.status, np.asarray(r.X_rows))
    finally:
        _cabi.gp_set_option(b'nr3_kernel', 0)
        _cabi.gp_set_option(b'nr3_ring', -1)
        _cabi.gp_set_option(b'nr3_prefetch', -1)
    for key in ((0, 2), (1, 0)):
        for a, b in zip(out[0, 0], out[key]):
            assert np.array_equal(a, b), key
    assert (out[0, 0][1] == 0).all()

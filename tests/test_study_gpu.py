"""GPU tests of the study kernels and drivers: device-side synthesis of
the C4 scenario rows is bit-identical to the NumPy definition, the |V| summary is exact, and a
DerStudy equals solving the same rows as one explicit batch."""
import os
import sys

import numpy as np
import pytest

from helpers import REPO, feeder_path, rel_err

pytestmark = pytest.mark.gpu
sys.path.insert(0, REPO)


def _c4_tables(n_srows, H=8760, D=128):
    """Hour and DER tables of BASELINE config C4 exactly as bench.py:make_der_rows draws them."""
    h_all = np.arange(H)
    rng = np.random.Generator(np.random.PCG64(123))
    m_h = np.clip(0.6 + 0.25 * np.sin(2 * np.pi * ((h_all % 24) - 9) / 24) + 0.1 * np.sin(2 * np.pi * h_all / H)
                  + rng.normal(0, 0.02, H), 0.3, 1.2)
    pv_h = np.maximum(0.0, np.sin(np.pi * ((h_all % 24) - 6) / 12))
    mask, pen = np.zeros((D, n_srows)), np.zeros(D)
    for dd in range(D):
        r = np.random.Generator(np.random.PCG64(1230 + dd))
        frac = r.uniform(0.1, 0.4)
        mask[dd, r.permutation(n_srows)[:max(1, int(round(frac * n_srows)))]] = 1.0
        pen[dd] = r.uniform(0.0, 0.8)
    return m_h, pv_h, mask, pen


def _fbs123():
    from gigpower_b200.solution_fbs import SolutionFBS
    return SolutionFBS(os.path.join(REPO, 'gigpower_b200', 'feeders', 'ieee123_allwye_noxfm_noreg.dss'))


@pytest.mark.parametrize('first,S', [(0, 1000), (8760 * 3 - 17, 4099), (8760 * 128 - 5, 64), (123456, 1)])
def test_der_rows_are_bit_identical_to_the_numpy_definition(first, S):
    import bench
    from gigpower_b200.study import DerStudy
    sol = _fbs123()
    m_h, pv_h, mask, pen = _c4_tables(sol.tables.n_srows)
    study = DerStudy(sol, m_h, pv_h, mask, pen)
    got = study.spu_rows(first, S).cpu().numpy()
    want = bench.make_der_rows(sol.load_weights().sum(axis=1), S, first)
    assert got.shape == want.shape
    assert np.array_equal(got.real, want.real) and np.array_equal(got.imag, want.imag)


def test_vmag_summary_is_exact():
    import torch
    from gigpower_b200.study import voltage_summary
    rng = np.random.Generator(np.random.PCG64(3))
    V = rng.uniform(0.9, 1.1, size=(37, 5001))
    V[5, 17] = np.nan
    V[:, 100] = 1.0
    lo, hi, cnt = voltage_summary(V, vmin=0.95, vmax=1.05)
    lo, hi, cnt = lo.cpu().numpy(), hi.cpu().numpy(), cnt.cpu().numpy()
    ok = np.ones(V.shape[1], bool)
    ok[17] = False
    assert np.array_equal(lo[ok], V.min(axis=0)[ok]) and np.array_equal(hi[ok], V.max(axis=0)[ok])
    assert np.array_equal(cnt, (~((V >= 0.95) & (V <= 1.05))).sum(axis=0))
    assert cnt[100] == 0 and cnt[17] >= 1
    # a sub-range of a wider pitch
    lo2, _, cnt2 = voltage_summary(torch.from_numpy(V).cuda(), S=1234)
    assert np.array_equal(lo2.cpu().numpy()[:17], lo[:17]) and np.array_equal(cnt2.cpu().numpy(), cnt[:1234])


@pytest.mark.parametrize('solver', ['fbs', 'nr3'])
def test_der_study_equals_the_explicit_batch(solver):
    import bench
    from gigpower_b200.study import DerStudy
    if solver == 'fbs':
        sol = _fbs123()
    else:
        from gigpower_b200.solution_nr3 import SolutionNR3
        sol = SolutionNR3(os.path.join(REPO, 'gigpower_b200', 'feeders', 'ieee123_allwye_noxfm_noreg.dss'))
    m_h, pv_h, mask, pen = _c4_tables(sol.tables.n_srows)
    first, count = 8760 * 5 + 100, 3000
    out = DerStudy(sol, m_h, pv_h, mask, pen).run(first=first, count=count, chunk=1024)
    rows = bench.make_der_rows(sol.load_weights().sum(axis=1), count, first)
    r = sol.solve_batch(spu_rows=rows, outputs=('Vmag', 'loss'))
    assert np.array_equal(out['iterations'], r.iterations) and np.array_equal(out['status'], r.status)
    assert (out['status'] == 0).all()
    assert rel_err(out['loss'], r.losses) <= 1e-13
    Vm = np.asarray(r.Vmag_rows)
    assert rel_err(out['vmin'], Vm.min(axis=0)) <= 1e-13 and rel_err(out['vmax'], Vm.max(axis=0)) <= 1e-13
    assert np.array_equal(out['violations'], (~((Vm >= 0.95) & (Vm <= 1.05))).sum(axis=0))
    assert out['vmin'].min() > 0.8 and out['vmax'].max() < 1.1


def test_scenario_frame_matches_the_snapshot_accessor():
    from gigpower_b200.solution_fbs import SolutionFBS
    from gigpower_b200.solution_nr3 import SolutionNR3
    from gigpower_b200.study import scenario_frame
    for cls in (SolutionFBS, SolutionNR3):
        s = cls(feeder_path('IEEE_13_Bus_allwye_noxfm_noreg'))
        s.solve()
        mult = np.ones((3, s.circuit.loads.num_elements))
        mult[1] = 0.7
        r = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag'))
        for param in ('V', 'Vmag', 'I'):
            snap = s.get_data_frame(param)
            fr = scenario_frame(s, r, param, 0)
            assert list(fr.index) == list(snap.index) and list(fr.columns) == list(snap.columns)
            assert rel_err(fr.values, snap.values) <= 1e-12
            for orient in ('rows', 'cols'):
                assert scenario_frame(s, r, param, 2, orient=orient).shape == s.get_data_frame(param, orient=orient).shape
        assert rel_err(scenario_frame(s, r, 'Vmag', 1).values, s.get_data_frame('Vmag').values) > 1e-4
        with pytest.raises(KeyError):
            scenario_frame(s, r, 'nope', 0)


def test_chained_study_matches_the_warm_started_oracle():
    """DerStudy.run_chained (hour k of a chain starts from hour k - 1) against the oracle driven the
    way a reused reference object runs (oracle/make_golden_warm.py pins that to the reference)."""
    import bench
    from gigpower_b200.dss_reader import read_dss
    from gigpower_b200.study import DerStudy
    from oracle.gp_oracle import FBSOracle
    sol = _fbs123()
    t = sol.tables
    m_h, pv_h, mask, pen = _c4_tables(t.n_srows)
    study = DerStudy(sol, m_h, pv_h, mask, pen)
    L, first_chain, chains = 24, 365 * 7 + 100, 20
    out = study.run_chained(chain_hours=L, first_chain=first_chain, chains=chains)
    o = FBSOracle(read_dss(os.path.join(REPO, 'gigpower_b200', 'feeders', 'ieee123_allwye_noxfm_noreg.dss')))
    W0 = sol.load_weights().sum(axis=1)
    b, p = np.nonzero(t.srow >= 0)
    V0 = I0 = None
    its, loss = np.zeros((chains, L), np.int64), np.zeros((chains, L), complex)
    for k in range(L):
        rows = np.concatenate([bench.make_der_rows(W0, 1, (first_chain + c) * L + k) for c in range(chains)], axis=1)
        spu = np.zeros((chains, o.c.nb, 3), complex)
        spu[:, b, p] = rows[t.srow[b, p]].T
        r = o.solve(spu, V0=V0 if k else None, I0=I0 if k else None)
        V0, I0 = r['V'], r['I']
        its[:, k], loss[:, k] = r['iterations'], (r['Stx'] - r['Srx']).sum(axis=(1, 2))
    assert np.array_equal(out['iterations'].reshape(chains, L), its)
    assert (out['status'] == 0).all()
    assert rel_err(out['loss'].reshape(chains, L), loss) <= 1e-9
    # against the flat-started study: same solutions within the solver tolerance, fewer sweeps
    flat = study.run(first=first_chain * L, count=chains * L)
    assert np.abs(out['vmin'] - flat['vmin']).max() <= 1e-5
    assert out['iterations'].sum() < flat['iterations'].sum()
    assert np.array_equal(out['iterations'].reshape(chains, L)[:, 0], flat['iterations'].reshape(chains, L)[:, 0])


def test_chained_study_nr3_agrees_with_the_flat_study():
    from gigpower_b200.solution_nr3 import SolutionNR3
    from gigpower_b200.study import DerStudy
    sol = SolutionNR3(os.path.join(REPO, 'gigpower_b200', 'feeders', 'ieee123_allwye_noxfm_noreg.dss'))
    m_h, pv_h, mask, pen = _c4_tables(sol.tables.n_srows)
    study = DerStudy(sol, m_h, pv_h, mask, pen)
    out = study.run_chained(chain_hours=12, first_chain=4000, chains=64)
    flat = study.run(first=4000 * 12, count=64 * 12)
    assert (out['status'] == 0).all()
    assert np.abs(out['vmin'] - flat['vmin']).max() <= 1e-8 and rel_err(out['loss'], flat['loss']) <= 1e-8
    assert out['iterations'].sum() < flat['iterations'].sum()
    with pytest.raises(ValueError):
        study.run_chained(chain_hours=7)

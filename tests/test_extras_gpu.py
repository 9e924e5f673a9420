"""GPU parity tests of the optional per-scenario inputs (GpSolveExtras, through the C-ABI):
regulator taps as a batch input and warm-started solves (8(f)-3).

Golden vectors come from the unmodified reference (oracle/make_golden_taps.py: one reference object
per scenario, built with that scenario's taps; oracle/make_golden_warm.py: one reference object
driven through a load sequence).  Bar: 1e-9 relative on complex voltages, identical iteration counts."""
import ctypes as C

import numpy as np
import pytest

from helpers import (ZIPS, LINES_ONLY, TAP_FEEDERS, WARM_FEEDERS, feeder_path, load_model, taps_golden, warm_golden,
                     rel_err)

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _fbs(feeder, **kw):
    from gigpower_b200.solution_fbs import SolutionFBS
    return SolutionFBS(feeder_path(feeder), zip_v=np.asarray(ZIPS['test_zip']), **kw)


def _nr3(feeder, **kw):
    from gigpower_b200.solution_nr3 import SolutionNR3
    return SolutionNR3(feeder_path(feeder), zip_v=np.asarray(ZIPS['test_zip']), **kw)


# ---------------------------------------------------------------------------------------------
# per-scenario taps
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('feeder', TAP_FEEDERS)
def test_fbs_taps_match_reference_objects(feeder):
    tg = taps_golden()
    s = _fbs(feeder)
    assert [vr.__name__ for vr in s.circuit.voltage_regulators.get_elements()] == list(tg[feeder + '/regs'])
    r = s.solve_batch(load_mult=tg[feeder + '/mult'], taps=tg[feeder + '/taps'], outputs=('V', 'I'))
    assert list(r.iterations) == list(tg[feeder + '/FBS/iterations'])
    assert rel_err(r.V, tg[feeder + '/FBS/V']) <= TOL
    # the reference keeps all-zero I rows for regulators after the line and transformer rows
    want_I = tg[feeder + '/FBS/I']
    assert rel_err(r.I, want_I[:, :r.I.shape[1]]) <= TOL


@pytest.mark.parametrize('feeder,S,n_combo,seed', [('IEEE_13_Bus_allwye', 700, 12, 11),
                                                    ('IEEE_34_Bus_allwye', 130, 6, 12),
                                                    ('IEEE_37_Bus_allwye', 257, 8, 13)])
def test_fbs_taps_random_batch_matches_oracle(feeder, S, n_combo, seed):
    from oracle.gp_oracle import FBSOracle, OracleCircuit, solve_with_taps
    rng = np.random.Generator(np.random.PCG64(seed))
    model = load_model(feeder)
    c = OracleCircuit(model)
    combos = rng.integers(-16, 17, size=(n_combo, len(model.regcontrols)))
    taps = combos[rng.integers(0, n_combo, size=S)]
    mult = rng.uniform(0.5, 1.5, size=(S, len(c.load_names)))
    want = solve_with_taps(FBSOracle, model, c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), taps, ZIPS['test_zip'])
    s = _fbs(feeder)
    r = s.solve_batch(load_mult=mult, taps=taps, outputs=('V', 'I', 'Vmag', 'loss'))
    assert np.array_equal(r.iterations, want['iterations'])
    assert np.array_equal(r.converged, want['converged'])
    assert rel_err(r.V, want['V']) <= TOL and rel_err(r.Vmag, want['Vmag']) <= TOL
    assert rel_err(r.losses, (want['Stx'] - want['Srx']).sum(axis=(1, 2))) <= TOL


def test_fbs_constant_taps_equal_a_feeder_built_with_them():
    """Ratios read per scenario or baked into the step table: the same sweeps (the two kernel instances
    may contract multiply-adds differently, so the comparison is at rounding level, not bitwise)."""
    from gigpower_b200 import _cabi
    feeder, taps = 'IEEE_13_Bus_allwye', {'reg1': -3, 'reg2': 12, 'reg3': 5}
    rng = np.random.Generator(np.random.PCG64(5))
    mult = rng.uniform(0.5, 1.5, size=(333, 15))
    s = _fbs(feeder)
    a = s.solve_batch(load_mult=mult, taps=taps, outputs=('V', 'I'))                 # dict form, scalars
    tv = np.tile(np.array([[-3, 12, 5]]), (333, 1))
    a2 = s.solve_batch(load_mult=mult, taps=tv, outputs=('V', 'I'))
    model = load_model(feeder)
    for rc in model.regcontrols:
        rc.tap = taps[rc.name]
    from gigpower_b200.solution_fbs import SolutionFBS
    b_sol = SolutionFBS(model, zip_v=np.asarray(ZIPS['test_zip']))
    b_sol._dev()
    _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
    try:
        b = b_sol.solve_batch(load_mult=mult, outputs=('V', 'I'))
    finally:
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 0))
    for r in (a, a2):
        assert np.array_equal(r.iterations, b.iterations)
        assert rel_err(r.V_rows, b.V_rows) <= 1e-13 and rel_err(r.I_rows, b.I_rows) <= 1e-13
    with pytest.raises(KeyError):
        s.solve_batch(load_mult=mult, taps={'nosuchreg': 1})
    with pytest.raises(ValueError):
        s.solve_batch(load_mult=mult, taps=tv[:, :2])


def test_taps_on_a_feeder_without_regulators_are_ignored():
    from gigpower_b200 import _cabi
    feeder = LINES_ONLY[0]
    s = _fbs(feeder)
    s._dev()
    assert _cabi.gp_gamma_rows(s._feeder.handle) == 0
    mult = np.random.Generator(np.random.PCG64(3)).uniform(0.5, 1.5, size=(100, s.circuit.loads.num_elements))
    a = s.solve_batch(load_mult=mult)
    b = s.solve_batch(load_mult=mult, taps=np.zeros((100, 0)))
    assert np.array_equal(a.iterations, b.iterations) and rel_err(a.V, b.V) <= 1e-12
    for f, make in (('IEEE_34_Bus_allwye', _fbs), ('IEEE_13_Bus_allwye', _nr3)):
        x = make(f)
        x._dev()
        assert _cabi.gp_gamma_rows(x._feeder.handle) == len(x.tables.grow_vr) > 0


@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye', 'IEEE_37_Bus_allwye'])
def test_nr3_taps_match_reference_objects(feeder):
    tg = taps_golden()
    X = tg[feeder + '/NR3/XNR']
    n = len(X)
    s = _nr3(feeder)
    r = s.solve_batch(load_mult=tg[feeder + '/mult'][:n], taps=tg[feeder + '/taps'][:n], outputs=('X', 'V'))
    assert list(r.iterations) == list(tg[feeder + '/NR3/iterations'])
    assert rel_err(r.XNR, X) <= TOL
    assert rel_err(r.V, tg[feeder + '/NR3/V']) <= TOL


@pytest.mark.parametrize('feeder,S,n_combo,seed', [('IEEE_13_Bus_allwye', 300, 10, 21), ('IEEE_37_Bus_allwye', 200, 6, 22),
                                                    ('IEEE_34_Bus_allwye', 150, 5, 23)])
def test_nr3_taps_random_batch_matches_tree_oracle(feeder, S, n_combo, seed):
    from oracle.gp_oracle import OracleCircuit, solve_with_taps
    from oracle.nr3_tree import NR3TreeOracle
    rng = np.random.Generator(np.random.PCG64(seed))
    model = load_model(feeder)
    c = OracleCircuit(model)
    combos = rng.integers(-16, 17, size=(n_combo, len(model.regcontrols)))
    taps = combos[rng.integers(0, n_combo, size=S)]
    mult = rng.uniform(0.5, 1.3, size=(S, len(c.load_names)))
    want = solve_with_taps(NR3TreeOracle, model, c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), taps,
                           ZIPS['test_zip'])
    r = _nr3(feeder).solve_batch(load_mult=mult, taps=taps, outputs=('X',))
    assert np.array_equal(r.iterations, want['iterations'])
    assert rel_err(r.XNR, want['XNR']) <= TOL


# ---------------------------------------------------------------------------------------------
# warm start
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('feeder', list(WARM_FEEDERS))
def test_fbs_warm_sequence_matches_reference_object(feeder):
    wg = warm_golden()
    s = _fbs(feeder, taps=WARM_FEEDERS[feeder])
    mult = wg[feeder + '/mult']
    prev = None
    for h in range(len(wg[feeder + '/FBS/iterations'])):
        prev = s.solve_batch(load_mult=mult[h:h + 1], outputs=('V', 'I'), return_device=True, warm_start=prev)
        assert int(prev.iterations[0]) == int(wg[feeder + '/FBS/iterations'][h]), h
        assert rel_err(prev.V[0], wg[feeder + '/FBS/V'][h]) <= TOL
        want_I = wg[feeder + '/FBS/I'][h]
        assert rel_err(prev.I[0], want_I[:prev.I.shape[1]]) <= TOL


@pytest.mark.parametrize('feeder,S,steps,seed', [('IEEE_13_Bus_allwye_noxfm_noreg', 1000, 4, 31),
                                                  ('IEEE_37_Bus_allwye', 150, 3, 32)])
def test_fbs_warm_batch_matches_oracle(feeder, S, steps, seed):
    from oracle.gp_oracle import FBSOracle
    rng = np.random.Generator(np.random.PCG64(seed))
    o = FBSOracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    s = _fbs(feeder)
    level = rng.uniform(0.6, 1.2, size=(S, 1))
    prev, V0, I0 = None, None, None
    total_warm = total_flat = 0
    for h in range(steps):
        level = level * rng.uniform(0.9, 1.1, size=(S, 1))
        mult = level * (1 + rng.normal(0, 0.02, size=(S, len(c.load_names))))
        want = o.solve(c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), V0=V0, I0=I0)
        prev = s.solve_batch(load_mult=mult, outputs=('V', 'I'), return_device=True, warm_start=prev)
        assert np.array_equal(prev.iterations.cpu().numpy(), want['iterations']), h
        assert rel_err(prev.V, want['V']) <= TOL
        V0, I0 = want['V'], want['I']
        if h > 0:
            total_warm += int(want['iterations'].sum())
            total_flat += int(s.solve_batch(load_mult=mult).iterations.sum())
    assert total_warm < total_flat          # the point of the warm start


@pytest.mark.parametrize('feeder', list(WARM_FEEDERS))
def test_nr3_warm_sequence_matches_reference_object(feeder):
    wg = warm_golden()
    s = _nr3(feeder, taps=WARM_FEEDERS[feeder])
    mult = wg[feeder + '/mult']
    prev = None
    for h in range(len(wg[feeder + '/NR3/iterations'])):
        prev = s.solve_batch(load_mult=mult[h:h + 1], outputs=('X',), return_device=True, warm_start=prev)
        assert int(prev.iterations[0]) == int(wg[feeder + '/NR3/iterations'][h]), h
        assert rel_err(prev.XNR[0], wg[feeder + '/NR3/XNR'][h]) <= TOL


def test_nr3_warm_batch_matches_dense_oracle():
    from oracle.gp_oracle import NR3Oracle
    feeder, S = 'IEEE_13_Bus_allwye_noxfm_noreg', 96
    rng = np.random.Generator(np.random.PCG64(41))
    o = NR3Oracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    s = _nr3(feeder)
    level = rng.uniform(0.6, 1.2, size=(S, 1))
    prev, X0 = None, None
    for h in range(3):
        level = level * rng.uniform(0.9, 1.1, size=(S, 1))
        mult = level * (1 + rng.normal(0, 0.02, size=(S, len(c.load_names))))
        want = o.solve(c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), X0=X0)
        prev = s.solve_batch(load_mult=mult, outputs=('X',), return_device=True, warm_start=prev)
        assert np.array_equal(prev.iterations.cpu().numpy(), want['iterations']), h
        assert rel_err(prev.XNR, want['XNR']) <= TOL
        X0 = want['XNR']


def test_extras_argument_checks():
    from gigpower_b200 import _cabi
    import torch
    s = _fbs('IEEE_13_Bus_allwye')
    _, dev = s._dev()
    t = s.tables
    S = 8
    spu = torch.zeros((t.n_srows, S), dtype=torch.complex128, device=dev)
    V = torch.zeros((t.n_vrows, S), dtype=torch.complex128, device=dev)
    I = torch.zeros((t.n_irows, S), dtype=torch.complex128, device=dev)
    it = torch.zeros(S, dtype=torch.int32, device=dev)
    st = torch.zeros(S, dtype=torch.int32, device=dev)
    df = torch.zeros(S, dtype=torch.float64, device=dev)
    ex = _cabi.GpSolveExtras()
    ex.warm_start = 2
    rc = _cabi.gp_fbs_solve_batch_ex(s._feeder.handle, S, S, spu.data_ptr(), C.byref(ex), V.data_ptr(), I.data_ptr(),
                                     it.data_ptr(), st.data_ptr(), df.data_ptr(), None, None, 1e-6, 100, None)
    assert rc == -1
    with pytest.raises(ValueError):
        s.solve_batch(load_mult=np.ones((4, 15)), warm_start=s.solve_batch(load_mult=np.ones((5, 15)),
                                                                           outputs=('V', 'I'), return_device=True))
    # extras = NULL is the plain entry point
    rc = _cabi.gp_fbs_solve_batch_ex(s._feeder.handle, S, S, spu.data_ptr(), None, V.data_ptr(), I.data_ptr(),
                                     it.data_ptr(), st.data_ptr(), df.data_ptr(), None, None, 1e-6, 100, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert int(it.max()) >= 1


# ---------------------------------------------------------------------------------------------
# Solution.wpu (reactive injection) and the volt-var droop
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('feeder', list(WARM_FEEDERS))
def test_fbs_reactive_injection_matches_reference_objects(feeder):
    from helpers import wpu_golden
    g = wpu_golden()
    s = _fbs(feeder, taps=WARM_FEEDERS[feeder])
    r = s.solve_batch(load_mult=g[feeder + '/mult'], wpu=g[feeder + '/wpu'], outputs=('V', 'I', 'sV'))
    assert list(r.iterations) == list(g[feeder + '/FBS/iterations'])
    assert rel_err(r.V, g[feeder + '/FBS/V']) <= TOL and rel_err(r.sV, g[feeder + '/FBS/sV']) <= TOL
    assert rel_err(r.I, g[feeder + '/FBS/I'][:, :r.I.shape[1]]) <= TOL
    # without the injection the voltages differ (the term is really applied)
    r0 = s.solve_batch(load_mult=g[feeder + '/mult'], outputs=('V',))
    assert rel_err(r0.V, g[feeder + '/FBS/V']) > 1e-5


def test_fbs_reactive_injection_random_batch_matches_oracle():
    from oracle.gp_oracle import FBSOracle
    feeder, S = 'IEEE_34_Bus_allwye', 400
    rng = np.random.Generator(np.random.PCG64(51))
    o = FBSOracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    mult = rng.uniform(0.5, 1.3, size=(S, len(c.load_names)))
    wpu = rng.uniform(-0.02, 0.02, size=(S, c.nb, 3)) * (c.bus_pm[None] != 0) * (rng.uniform(size=(S, c.nb, 3)) < 0.3)
    want = o.solve(c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), wpu=wpu)
    s = _fbs(feeder)
    r = s.solve_batch(load_mult=mult, wpu=wpu, outputs=('V', 'sV', 'Vmag', 'loss'))
    assert np.array_equal(r.iterations, want['iterations'])
    assert rel_err(r.V, want['V']) <= TOL and rel_err(r.sV, want['sV']) <= TOL
    # packed rows on the device give the same
    import torch
    b, p = np.nonzero(s.tables.vrow >= 0)
    rows = np.zeros((s.tables.n_vrows, S))
    rows[s.tables.vrow[b, p]] = wpu[:, b, p].T
    r2 = s.solve_batch(load_mult=mult, wpu=torch.from_numpy(rows).cuda(), outputs=('V',))
    assert np.array_equal(r2.V_rows, r.V_rows)
    with pytest.raises(ValueError):
        s.solve_batch(load_mult=mult, wpu=rows[:, :5])


@pytest.mark.parametrize('feeder', list(WARM_FEEDERS))
def test_volt_var_study_matches_the_reference_controller(feeder):
    from helpers import wpu_golden
    from gigpower_b200.study import VoltVarStudy
    g = wpu_golden()
    s = _fbs(feeder, taps=WARM_FEEDERS[feeder])
    want_V, want_Q, want_it = g[feeder + '/VVC/V'], g[feeder + '/VVC/Q'], g[feeder + '/VVC/iterations']
    n, rounds = want_V.shape[0], want_V.shape[1]
    minQ, maxQ = g[feeder + '/VVC/q']
    study = VoltVarStudy(s, [tuple(int(v) for v in x) for x in g[feeder + '/VVC/ctl']],
                         points=tuple(g[feeder + '/VVC/points']), minQ=minQ, maxQ=maxQ)
    last, Q, its = study.run(load_mult=g[feeder + '/mult'][:n], rounds=rounds)
    assert np.array_equal(its.cpu().numpy().T, want_it)
    assert rel_err(Q.cpu().numpy().transpose(2, 0, 1), want_Q) <= TOL
    assert rel_err(last.V, want_V[:, -1]) <= TOL


def test_volt_var_study_batch_matches_oracle_loop():
    from oracle.gp_oracle import FBSOracle, fbs_volt_var
    from gigpower_b200.study import VoltVarStudy
    feeder, S, rounds = 'IEEE_13_Bus_allwye_noxfm_noreg', 300, 5
    rng = np.random.Generator(np.random.PCG64(61))
    o = FBSOracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    mult = rng.uniform(0.4, 1.6, size=(S, len(c.load_names)))
    ctl = [(c.bus_names.index('675'), p) for p in range(3)] + [(c.bus_names.index('634'), 1)]
    pts = (0.93, 0.97, 1.0, 1.04)
    V, Q, its = fbs_volt_var(o, c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), ctl, pts, -0.1, 0.1, rounds)
    s = _fbs(feeder)
    study = VoltVarStudy(s, [('675', 0), ('675', 1), ('675', 2), ('634', 1)], points=pts, minQ=-0.1, maxQ=0.1)
    last, Qg, itg = study.run(load_mult=mult, rounds=rounds)
    assert np.array_equal(itg.cpu().numpy(), its)
    assert rel_err(Qg.cpu().numpy().transpose(0, 2, 1), Q) <= TOL
    assert rel_err(last.V, V[-1]) <= TOL
    assert np.abs(Q).max() > 0.01 and (its[-1] <= its[0]).all()


# ---------------------------------------------------------------------------------------------
# wpu in SolutionNR3 (change_KCL_matrices' "- wpu" and map_output's "+ 1j wpu")
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_13_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg'])
def test_nr3_reactive_injection_matches_reference_objects(feeder):
    """tests/golden/ref_wpu_nr3.npz: unmodified reference SolutionNR3 objects whose circuit reports a non-zero wpu
    matrix, after the reference's own change_KCL_matrices (oracle/make_golden_wpu_nr3.py)."""
    import os
    from helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'ref_wpu_nr3.npz'))
    s = _nr3(feeder, taps=WARM_FEEDERS.get(feeder, {}))
    r = s.solve_batch(load_mult=g[feeder + '/mult'], wpu=g[feeder + '/wpu'], outputs=('X', 'V', 'sV', 'i_Node'))
    assert list(r.iterations) == list(g[feeder + '/NR3/iterations'])
    assert rel_err(r.XNR, g[feeder + '/NR3/XNR']) <= TOL and rel_err(r.V, g[feeder + '/NR3/V']) <= TOL
    assert rel_err(r.sV, g[feeder + '/NR3/sV']) <= TOL and rel_err(r.i_Node, g[feeder + '/NR3/i_Node']) <= TOL
    r0 = s.solve_batch(load_mult=g[feeder + '/mult'], outputs=('X',))
    assert rel_err(r0.XNR, g[feeder + '/NR3/XNR']) > 1e-5       # the term is really applied


def test_nr3_reactive_injection_random_batch_matches_tree_oracle():
    import torch
    from oracle.nr3_tree import NR3TreeOracle
    feeder, S = 'IEEE_34_Bus_allwye_noxfm_noreg', 300
    rng = np.random.Generator(np.random.PCG64(52))
    o = NR3TreeOracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    mult = rng.uniform(0.5, 1.2, size=(S, len(c.load_names)))
    wpu = rng.uniform(-0.02, 0.02, size=(S, c.nb, 3)) * (c.bus_pm[None] != 0) * (rng.uniform(size=(S, c.nb, 3)) < 0.3)
    wpu[:, 0] = 0
    want = o.solve(c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), wpu=wpu)
    s = _nr3(feeder)
    r = s.solve_batch(load_mult=mult, wpu=wpu, outputs=('X',))
    assert np.array_equal(r.iterations, want['iterations'])
    assert rel_err(r.XNR, want['XNR']) <= TOL
    b, p = np.nonzero(s.tables.vrow >= 0)
    rows = np.zeros((s.tables.n_vrows, S))
    rows[s.tables.vrow[b, p]] = wpu[:, b, p].T
    r2 = s.solve_batch(load_mult=mult, wpu=torch.from_numpy(rows).cuda(), outputs=('X',))
    assert np.array_equal(np.asarray(r2.X_rows), np.asarray(r.X_rows))
    with pytest.raises(ValueError):
        s.solve_batch(load_mult=mult, wpu=rows[:, :5])


def test_volt_var_study_runs_on_nr3_too():
    """The quasi-static droop loop (solve, wpu = get_Q(|V|), solve again from the last X) around SolutionNR3:
    every round is held to the tree oracle given the same wpu, and the injections move the controlled voltages
    towards the dead band."""
    from oracle.nr3_tree import NR3TreeOracle
    from gigpower_b200.study import VoltVarStudy
    feeder, S, rounds = 'IEEE_13_Bus_allwye_noxfm_noreg', 64, 3
    rng = np.random.Generator(np.random.PCG64(62))
    o = NR3TreeOracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    mult = rng.uniform(0.8, 1.6, size=(S, len(c.load_names)))
    s = _nr3(feeder)
    pts = (0.93, 0.97, 1.0, 1.04)
    study = VoltVarStudy(s, [('675', 0), ('675', 1), ('675', 2)], points=pts, minQ=-0.1, maxQ=0.1)
    last, Q, its = study.run(load_mult=mult, rounds=rounds, outputs=('X', 'Vmag'))
    Q = Q.cpu().numpy()                                     # [rounds, 3, S]
    assert np.abs(Q).max() > 0.01
    # the last round's solve used wpu = Q[rounds - 2]: flat-started oracle with that wpu gives the same X (Newton
    # converges to the same solution from either start)
    k = c.bus_names.index('675')
    wpu = np.zeros((S, c.nb, 3))
    wpu[:, k, :] = Q[rounds - 2].T
    want = o.solve(c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), wpu=wpu)
    assert rel_err(last.XNR, want['XNR']) <= 1e-8


# ---------------------------------------------------------------------------------------------
# volt-var droop inside the FBS iteration (GpSolveExtras.vv_*)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_13_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg'])
def test_droop_inside_the_iteration_matches_the_reference_with_its_controllers(feeder):
    """tests/golden/ref_vvc_iter.npz (oracle/make_golden_vvc_iter.py): the reference's SolutionFBS with its own
    VoltVARController objects consulted before every calc_sV - including scenarios in which the droop keeps
    the sweep from converging within maxiter (equal status and iteration counts)."""
    import os
    from helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'ref_vvc_iter.npz'))
    s = _fbs(feeder, taps=WARM_FEEDERS.get(feeder, {}))
    ctl = [tuple(int(v) for v in x) for x in g[feeder + '/ctl']]
    minQ, maxQ = g['q']
    r = s.solve_batch(load_mult=g[feeder + '/mult'], outputs=('V', 'I', 'sV'),
                      volt_var=(ctl, tuple(g['points']), minQ, maxQ))
    want_it = g[feeder + '/iterations']
    assert list(r.iterations) == list(want_it)
    ok = want_it < 100
    assert (r.status[ok] == 0).all() and (r.status[~ok] != 0).all()
    assert rel_err(r.V, g[feeder + '/V']) <= TOL and rel_err(r.sV, g[feeder + '/sV']) <= TOL
    assert rel_err(r.I, g[feeder + '/I'][:, :r.I.shape[1]]) <= TOL
    r0 = s.solve_batch(load_mult=g[feeder + '/mult'], outputs=('V',))
    assert rel_err(r0.V, g[feeder + '/V']) > 1e-5


def test_droop_inside_the_iteration_random_batch_matches_oracle():
    from oracle.gp_oracle import FBSOracle
    from gigpower_b200.study import VoltVarStudy
    feeder, S = 'IEEE_34_Bus_allwye_noxfm_noreg', 500
    rng = np.random.Generator(np.random.PCG64(71))
    o = FBSOracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    mult = rng.uniform(0.4, 1.2, size=(S, len(c.load_names)))
    three = [k for k in range(1, c.nb) if c.bus_pm[k].all()]
    ctl = [(three[len(three) // 3], 0), (three[len(three) // 3], 2), (three[-1], 1), (three[-2], 0)]
    pts, minQ, maxQ = (0.92, 0.96, 1.0, 1.04), -0.03, 0.03
    want = o.solve(c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), droop=(ctl, pts, minQ, maxQ))
    s = _fbs(feeder)
    study = VoltVarStudy(s, ctl, points=pts, minQ=minQ, maxQ=maxQ)
    r, Q = study.run_in_iteration(load_mult=mult, outputs=('V', 'Vmag'))
    it, st = r.iterations.cpu().numpy(), r.status.cpu().numpy()
    assert np.array_equal(it, want['iterations'])
    ok = st == 0
    assert ok.sum() > S // 2
    assert rel_err(r.V[ok], want['V'][ok]) <= TOL
    assert float(Q.abs().max()) > 1e-3


# ---------------------------------------------------------------------------------------------
# per-scenario taps and warm starts on the feeder-specialised kernel (its EX build)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye', 'IEEE_37_Bus_allwye', 'IEEE_34_Bus_allwye',
                                    'IEEE_13_Bus_allwye_noxfm_noreg'])
def test_specialised_kernel_takes_taps_and_warm_starts(feeder):
    """Once gp_fbs_specialise has built the feeder's kernel, gp_fbs_solve_batch_ex with regulator ratios and / or a
    warm start runs on that kernel's EX build (one launch incl. |V| and losses) instead of the streaming kernel:
    same iteration counts and status, V / I / |V| / losses within 1e-12 of the streaming kernel's, on ragged batch
    sizes; the golden per-scenario-tap vectors of the reference are met by the taps tests above through either."""
    from gigpower_b200 import _cabi
    s = _fbs(feeder)
    S = 2111
    rng = np.random.Generator(np.random.PCG64(81))
    mult = rng.uniform(0.5, 1.3, size=(S, s.circuit.loads.num_elements))
    nreg = s.circuit.voltage_regulators.num_elements
    taps = rng.integers(-16, 17, size=(S, nreg)) if nreg else None
    outs = ('V', 'I', 'Vmag', 'loss')

    def run():
        a = s.solve_batch(load_mult=mult, taps=taps, outputs=outs, return_device=True)
        n0 = _cabi.gp_launch_count()
        b = s.solve_batch(load_mult=mult * 1.07, taps=taps, outputs=outs, warm_start=a)
        return a, b, _cabi.gp_launch_count() - n0

    try:
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 3))
        a3, b3, n3 = run()
        a3 = (a3.iterations.cpu().numpy(), a3.status.cpu().numpy())
        _cabi.check(_cabi.gp_set_option(b'fbs_kernel', 0))
        assert s.specialise()
        a5, b5, n5 = run()
        assert n5 == 1 and n3 >= 2                      # one fused launch against solve + |V| + loss kernels
        assert np.array_equal(a5.iterations.cpu().numpy(), a3[0]) and np.array_equal(a5.status.cpu().numpy(), a3[1])
        assert np.array_equal(b5.iterations, b3.iterations) and np.array_equal(b5.status, b3.status)
        assert (b3.status == 0).all() and b3.iterations.mean() < a3[0].mean()      # the warm start is a warm start
        assert rel_err(b5.V, b3.V) <= 1e-12 and rel_err(b5.I, b3.I) <= 1e-12
        assert rel_err(b5.Vmag, b3.Vmag) <= 1e-12 and rel_err(b5.losses, b3.losses) <= 1e-12
    finally:
        _cabi.gp_set_option(b'fbs_kernel', 0)

"""Pin the CPU oracle (oracle/gp_oracle.py) to the reference.

* full-precision vectors produced by the unmodified reference run behind the
  opendssdirect stand-in (oracle/make_golden.py -> tests/golden/*.npz);
* gigpower's own recorded result tables (two decimals), parsed from
  gigpower/tests/test_compare_{fbs,nr3}_dss/*.out.txt;
* the known-answer values of SURVEY.md appendix B.
"""
import numpy as np
import pytest

from oracle.gp_oracle import FBSOracle, NR3Oracle, OracleCircuit
from helpers import (ZIPS, ALL_FEEDERS, LINES_ONLY, feeder_path, load_model, snapshots, scenarios, recorded,
                     pinned_34, pinned_34_taps)

TOL = 1e-12      # oracle vs reference, absolute on pu quantities


@pytest.mark.parametrize('zname', list(ZIPS))
@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_fbs_snapshot(feeder, zname):
    snap = snapshots()
    o = FBSOracle(load_model(feeder), ZIPS[zname])
    r = o.solve()
    k = f'{feeder}/{zname}/FBS/'
    assert o.topo_order == list(snap[k + 'topo_order'])
    assert int(r['iterations'][0]) == int(snap[k + 'iterations'])
    assert bool(r['converged'][0]) == bool(snap[k + 'converged'])
    assert abs(r['convergence_diff'][0] - float(snap[k + 'convergence_diff'])) <= 1e-15
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx'):
        assert np.abs(r[p][0] - snap[k + p]).max() <= TOL, p


@pytest.mark.parametrize('zname', list(ZIPS))
@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_nr3_snapshot(feeder, zname):
    snap = snapshots()
    r = NR3Oracle(load_model(feeder), ZIPS[zname]).solve()
    k = f'{feeder}/{zname}/NR3/'
    assert int(r['iterations'][0]) == int(snap[k + 'iterations'])
    assert np.abs(r['XNR'][0] - snap[k + 'XNR'][:, 0]).max() <= TOL
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx', 'i_Node'):
        assert np.abs(r[p][0] - snap[k + p]).max() <= TOL, p


@pytest.mark.parametrize('feeder', LINES_ONLY)
def test_fbs_scenarios(feeder):
    sc = scenarios()
    o = FBSOracle(load_model(feeder), ZIPS['test_zip'])
    mult = sc[feeder + '/mult']
    c = o.c
    spu = c.spu_matrix(c.load_kw * mult, c.load_kvar * mult)
    r = o.solve(spu)
    assert list(r['iterations']) == list(sc[feeder + '/FBS/iterations'])
    assert np.abs(r['V'] - sc[feeder + '/FBS/V']).max() <= TOL
    assert np.abs(r['I'] - sc[feeder + '/FBS/I']).max() <= TOL


@pytest.mark.parametrize('feeder', LINES_ONLY)
def test_nr3_scenarios(feeder):
    sc = scenarios()
    o = NR3Oracle(load_model(feeder), ZIPS['test_zip'])
    X = sc[feeder + '/NR3/XNR']
    mult = sc[feeder + '/mult'][:len(X)]
    c = o.c
    r = o.solve(c.spu_matrix(c.load_kw * mult, c.load_kvar * mult))
    assert list(r['iterations']) == list(sc[feeder + '/NR3/iterations'])
    assert np.abs(r['XNR'] - X).max() <= 1e-11


# The 34-bus feeder with regulators has its controls ON in OpenDSS and the tap
# positions OpenDSS chose are recorded nowhere in the reference; they are recovered
# from the tables themselves by oracle/pin_taps_34.py (unique within the search:
# every neighbouring tap vector misses >= 22 entries of a ZIP set).  The feeder
# fixture keeps taps 0 (ref_snapshots.npz pins that state); the tables need the
# pinned taps.
def _model(feeder):
    if feeder == 'IEEE_34_Bus_allwye':
        from gigpower_b200.dss_reader import read_dss
        return read_dss(feeder_path(feeder), pinned_34_taps())
    return load_model(feeder)


@pytest.mark.parametrize('algo', ['FBS', 'NR3'])
@pytest.mark.parametrize('zname', list(ZIPS))
@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_recorded_tables(feeder, zname, algo):
    """gigpower's recorded tables print '{:.2f}': agreement within rounding."""
    rec = recorded()
    model = _model(feeder)
    if algo == 'FBS':
        o = FBSOracle(model, ZIPS[zname])
        r = {k: v[0] for k, v in o.solve().items() if k in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx')}
        r['I'], names_l = r['I'][:o.c.nl], o.c.line_names
    else:
        o = NR3Oracle(model, ZIPS[zname])
        r = {k: v[0].T for k, v in o.solve().items() if k in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx')}
        names_l = o.c.line_names
    n = 0
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx'):
        table = rec[f'{feeder}/{zname}/{algo}/{p}']
        names = o.c.bus_names if p in ('V', 'sV', 'Vmag') else names_l
        assert list(table) == list(names)
        for i, name in enumerate(names):
            want = np.array([complex(*x) for x in table[name]])
            assert np.abs(r[p][i].real - want.real).max() <= 0.005 + 1e-9, (p, name)
            assert np.abs(r[p][i].imag - want.imag).max() <= 0.005 + 1e-9, (p, name)
            n += 3
    assert n > 0


def test_pinned_34_bus_taps():
    assert pinned_34_taps() == {'reg1a': 14, 'reg1b': 4, 'reg1c': 6, 'reg2a': 13, 'reg2b': 12, 'reg2c': 14}


@pytest.mark.parametrize('zname', list(ZIPS))
def test_34_bus_at_the_pinned_taps_matches_the_reference(zname):
    """Full precision: the unmodified reference at the taps of its recorded tables."""
    z = pinned_34()
    model = _model('IEEE_34_Bus_allwye')
    r = FBSOracle(model, ZIPS[zname]).solve()
    k = f'{zname}/FBS/'
    assert int(r['iterations'][0]) == int(z[k + 'iterations'])
    assert abs(float(r['convergence_diff'][0]) - float(z[k + 'convergence_diff'])) <= TOL
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx'):
        assert r[p][0].shape == z[k + p].shape, p
        assert np.abs(r[p][0] - z[k + p]).max() <= TOL, p
    r = NR3Oracle(model, ZIPS[zname]).solve()
    k = f'{zname}/NR3/'
    assert int(r['iterations'][0]) == int(z[k + 'iterations'])
    assert np.abs(r['XNR'][0] - z[k + 'XNR'][:, 0]).max() <= 1e-11
    for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx', 'i_Node'):
        assert np.abs(r[p][0] - z[k + p]).max() <= 1e-11, p


def test_known_answers_appendix_b():
    """values of the reference on the 13-bus lines-only feeder."""
    model = load_model('IEEE_13_Bus_allwye_noxfm_noreg')
    c = OracleCircuit(model)
    l = c.line_names.index('line_650_632')
    assert abs(c.zbase[c.line_tx[l]] - 5.768533333333334) < 1e-15
    assert abs(c.FZpu[l][0, 0] - (0.022752750554733723 + 0.06683989838286714j)) < 1e-17
    r = FBSOracle(model).solve()
    b = c.bus_names.index('675')
    want = np.array([0.9085896347716701 - 0.11171348099844346j, -0.5405817165643256 - 0.8531996289341933j,
                     -0.3838855973406417 + 0.7987657480063847j])
    assert np.abs(r['V'][0, b] - want).max() < 1e-14
    assert int(r['iterations'][0]) == 8
    assert abs(r['convergence_diff'][0] - 3.2439829541267554e-07) < 1e-18
    rn = NR3Oracle(model).solve()
    want = np.array([0.9186175866955602 - 0.10464543644259171j, -0.5360926994597318 - 0.8502195082139851j,
                     -0.3880887632230724 + 0.8105221042526009j])
    assert np.abs(rn['V'][0, :, b] - want).max() < 1e-13


@pytest.mark.parametrize('zname', list(ZIPS))
@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_tree_elimination_oracle_equals_dense_oracle(feeder, zname):
    """oracle/nr3_tree.py (the algorithm the CUDA kernel implements) against the
    literal restatement (dense J, inv(J'J) J' F) and the reference's golden X."""
    from oracle.nr3_tree import NR3TreeOracle
    model = load_model(feeder)
    o, t = NR3Oracle(model, ZIPS[zname]), NR3TreeOracle(model, ZIPS[zname])
    rng = np.random.Generator(np.random.PCG64(5))
    mult = rng.uniform(0.5, 1.3, size=(5, len(o.c.load_names)))
    mult[0] = 1.0
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    a, b = o.solve(spu), t.solve(spu)
    assert np.array_equal(a['iterations'], b['iterations'])
    assert np.abs(a['XNR'] - b['XNR']).max() <= 1e-12
    assert np.abs(b['XNR'][0] - snapshots()[f'{feeder}/{zname}/NR3/XNR'][:, 0]).max() <= 1e-12


# ---- per-scenario regulator taps --------------------------------
from oracle.gp_oracle import solve_with_taps, gamma_of_tap          # noqa: E402
from helpers import taps_golden, TAP_FEEDERS                         # noqa: E402


def test_gamma_of_tap_is_the_reference_mapping():
    # voltage_regulator.py:23-39: -16 -> 0.90, +16 -> 1.10, and the rounding of the python expression
    for tap in range(-16, 17):
        g = (tap + 16) / 32
        g *= 1.10 - .90
        g += .90
        assert gamma_of_tap(tap) == g
    assert gamma_of_tap(np.array([-16, 16]))[0] == .90 and abs(gamma_of_tap(16) - 1.10) < 1e-15


@pytest.mark.parametrize('feeder', TAP_FEEDERS)
def test_fbs_per_scenario_taps(feeder):
    tg = taps_golden()
    model = load_model(feeder)
    assert [rc.name for rc in model.regcontrols] == list(tg[feeder + '/regs'])
    c = OracleCircuit(model)
    mult = tg[feeder + '/mult']
    spu = c.spu_matrix(c.load_kw * mult, c.load_kvar * mult)
    r = solve_with_taps(FBSOracle, model, spu, tg[feeder + '/taps'], ZIPS['test_zip'])
    assert list(r['iterations']) == list(tg[feeder + '/FBS/iterations'])
    assert np.abs(r['V'] - tg[feeder + '/FBS/V']).max() <= TOL
    assert np.abs(r['I'] - tg[feeder + '/FBS/I']).max() <= TOL


@pytest.mark.parametrize('feeder', [f for f in TAP_FEEDERS if f != 'IEEE_34_Bus_allwye'])
def test_nr3_per_scenario_taps(feeder):
    from oracle.nr3_tree import NR3TreeOracle
    tg = taps_golden()
    model = load_model(feeder)
    c = OracleCircuit(model)
    X = tg[feeder + '/NR3/XNR']
    n = len(X)
    mult = tg[feeder + '/mult'][:n]
    spu = c.spu_matrix(c.load_kw * mult, c.load_kvar * mult)
    r = solve_with_taps(NR3Oracle, model, spu, tg[feeder + '/taps'][:n], ZIPS['test_zip'])
    assert list(r['iterations']) == list(tg[feeder + '/NR3/iterations'])
    assert np.abs(r['XNR'] - X).max() <= 1e-11
    t = solve_with_taps(NR3TreeOracle, model, spu, tg[feeder + '/taps'][:n], ZIPS['test_zip'])
    assert list(t['iterations']) == list(tg[feeder + '/NR3/iterations'])
    assert np.abs(t['XNR'] - X).max() <= 1e-10


# ---- warm-started sequences ---------------------------------------
from helpers import warm_golden, WARM_FEEDERS                       # noqa: E402


@pytest.mark.parametrize('feeder', list(WARM_FEEDERS))
def test_fbs_warm_started_sequence(feeder):
    wg = warm_golden()
    o = FBSOracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    mult = wg[feeder + '/mult']
    V0 = I0 = None
    for h in range(len(wg[feeder + '/FBS/iterations'])):
        r = o.solve(c.spu_matrix(c.load_kw * mult[h], c.load_kvar * mult[h])[None], V0=V0, I0=I0)
        assert int(r['iterations'][0]) == int(wg[feeder + '/FBS/iterations'][h]), h
        assert np.abs(r['V'][0] - wg[feeder + '/FBS/V'][h]).max() <= TOL
        assert np.abs(r['I'][0] - wg[feeder + '/FBS/I'][h]).max() <= TOL
        V0, I0 = r['V'], r['I']
    # the warm start is what saves iterations: the same loads from flat start take more sweeps
    flat = o.solve(c.spu_matrix(c.load_kw * mult[1], c.load_kvar * mult[1])[None])
    assert int(flat['iterations'][0]) > int(wg[feeder + '/FBS/iterations'][1])


@pytest.mark.parametrize('feeder', list(WARM_FEEDERS))
def test_nr3_warm_started_sequence(feeder):
    wg = warm_golden()
    o = NR3Oracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    mult = wg[feeder + '/mult']
    X0 = None
    for h in range(len(wg[feeder + '/NR3/iterations'])):
        r = o.solve(c.spu_matrix(c.load_kw * mult[h], c.load_kvar * mult[h])[None], X0=X0)
        assert int(r['iterations'][0]) == int(wg[feeder + '/NR3/iterations'][h]), h
        assert np.abs(r['XNR'][0] - wg[feeder + '/NR3/XNR'][h]).max() <= 1e-11
        X0 = r['XNR']


# ---- Solution.wpu and the volt-var droop -----------------------------
from oracle.gp_oracle import fbs_volt_var, vvc_get_Q                # noqa: E402
from helpers import wpu_golden                                       # noqa: E402


@pytest.mark.parametrize('feeder', list(WARM_FEEDERS))
def test_fbs_with_reactive_injection(feeder):
    g = wpu_golden()
    o = FBSOracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    mult = g[feeder + '/mult']
    r = o.solve(c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), wpu=g[feeder + '/wpu'])
    assert list(r['iterations']) == list(g[feeder + '/FBS/iterations'])
    for p in ('V', 'I', 'sV'):
        assert np.abs(r[p] - g[feeder + '/FBS/' + p]).max() <= TOL, p


@pytest.mark.parametrize('feeder', list(WARM_FEEDERS))
def test_volt_var_loop_matches_the_reference_controller(feeder):
    g = wpu_golden()
    o = FBSOracle(load_model(feeder), ZIPS['test_zip'])
    c = o.c
    n = g[feeder + '/VVC/V'].shape[0]
    mult = g[feeder + '/mult'][:n]
    ctl = [tuple(x) for x in g[feeder + '/VVC/ctl']]
    minQ, maxQ = g[feeder + '/VVC/q']
    V, Q, its = fbs_volt_var(o, c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), ctl, tuple(g[feeder + '/VVC/points']),
                             minQ, maxQ, g[feeder + '/VVC/V'].shape[1])
    assert np.array_equal(its.T, g[feeder + '/VVC/iterations'])
    assert np.abs(np.swapaxes(V, 0, 1) - g[feeder + '/VVC/V']).max() <= TOL
    assert np.abs(np.swapaxes(Q, 0, 1) - g[feeder + '/VVC/Q']).max() <= TOL
    assert np.abs(Q).max() > 1e-3            # the droop is active in these cases
    assert vvc_get_Q(np.array([0.9, 1.0, 1.1]), (0.95, 0.98, 1.02, 1.05), -0.05, 0.05).tolist() == [-0.05, -0.0, 0.05]


@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_13_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg'])
def test_nr3_tree_oracle_with_wpu_matches_the_reference(feeder):
    """tests/golden/ref_wpu_nr3.npz (oracle/make_golden_wpu_nr3.py): unmodified reference SolutionNR3 objects whose
    circuit reports a non-zero wpu matrix, after the reference's own change_KCL_matrices has written ``- wpu``
    into the constant term of their KCL rows.  The tree oracle's ``wpu`` argument must reproduce them, and the
    injection must matter (the same loads without it give a different answer)."""
    import os
    from helpers import GOLDEN, load_model, ZIPS
    from oracle.nr3_tree import NR3TreeOracle
    g = np.load(os.path.join(GOLDEN, 'ref_wpu_nr3.npz'))
    o = NR3TreeOracle(load_model(feeder), ZIPS['test_zip'])
    mult, wpu = g[feeder + '/mult'], g[feeder + '/wpu']
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    r = o.solve(spu, wpu=wpu)
    assert list(r['iterations']) == list(g[feeder + '/NR3/iterations'])
    assert np.abs(r['XNR'] - g[feeder + '/NR3/XNR']).max() <= 1e-11
    r0 = o.solve(spu)
    assert np.abs(r0['XNR'] - g[feeder + '/NR3/XNR']).max() >= 1e-5


@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_13_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg'])
def test_fbs_oracle_with_droop_inside_the_iteration_matches_the_reference(feeder):
    """tests/golden/ref_vvc_iter.npz (oracle/make_golden_vvc_iter.py): the reference's SolutionFBS with its own
    VoltVARController objects consulted before every calc_sV.  The oracle's ``droop`` argument reproduces V, I, sV,
    the iteration counts and the final injections."""
    import os
    from helpers import GOLDEN, load_model, ZIPS
    from oracle.gp_oracle import FBSOracle, vvc_get_Q
    g = np.load(os.path.join(GOLDEN, 'ref_vvc_iter.npz'))
    o = FBSOracle(load_model(feeder), ZIPS['test_zip'])
    mult = g[feeder + '/mult']
    ctl = [tuple(int(v) for v in x) for x in g[feeder + '/ctl']]
    minQ, maxQ = g['q']
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    r = o.solve(spu, droop=(ctl, tuple(g['points']), minQ, maxQ))
    assert list(r['iterations']) == list(g[feeder + '/iterations'])
    for k in ('V', 'I', 'sV'):
        assert np.abs(r[k] - g[feeder + '/' + k]).max() <= 1e-12, k
    Q = np.stack([vvc_get_Q(np.abs(r['V'][:, k, p]), tuple(g['points']), minQ, maxQ) for k, p in ctl], axis=1)
    assert np.abs(Q - g[feeder + '/Q']).max() <= 1e-12 and np.abs(Q).max() > 1e-3

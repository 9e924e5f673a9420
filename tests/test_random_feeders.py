"""Sixteen seeded RANDOM radial feeders (oracle/make_golden_random.py): random tree shapes, 1- / 2- / 3-phase laterals
(some declared in reverse phase order), four kinds of line definition, 1- / 2- / 3-phase loads, capacitors, per-phase
regulators with random taps and a 4.16 / 0.48 kV transformer on some.  The golden vectors are the UNMODIFIED
reference's own ``SolutionFBS`` / ``SolutionNR3`` results on them (tests/golden/ref_random.npz), so every check here
is against the reference, not against the oracle:

* the oracles (FBS; NR3 dense and tree elimination) - pins them on topologies nobody hand-picked;
* the streaming kernel's loop run in NumPy over the library's own schedule (``gp_fbs_schedule``), both step orders,
  and the tree-walk kernel's loop over ``gp_fbs_walk_events``;
* the GENERATED sweeps of the feeder-specialised kernel (two-sweep with I slots in shared or global memory, and the
  depth-first build), compiled for the host.

No GPU needed; the GPU counterpart is tests/test_random_feeders_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import GOLDEN, ZIPS
from gigpower_b200.dss_reader import read_dss

RANDOM = [f'rnd{i:02d}' for i in range(16)]
HOST_COMPILED = RANDOM
TOL = 1e-12
_cache = {}


def gold():
    if 'g' not in _cache:
        _cache['g'] = np.load(os.path.join(GOLDEN, 'ref_random.npz'))
    return _cache['g']


def rmodel(name):
    return read_dss(os.path.join(GOLDEN, 'random_feeders', name + '.json'))


def _tables(name):
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.flatten import flatten_fbs
    from gigpower_b200.solution import Solution
    Solution.set_zip_values(np.asarray(ZIPS['test_zip']))
    model = rmodel(name)
    return model, flatten_fbs(Circuit(model))


def _packed(t, o, mult):
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    b, p = np.nonzero(t.srow >= 0)
    rows = np.zeros((max(t.n_srows, 1), len(mult)), dtype=complex)
    rows[t.srow[b, p]] = spu[:, b, p].T
    return rows


def _check_rows(t, V, I, its, name, tol=TOL, i_rows=None):
    g = gold()
    assert list(its) == list(g[name + '/FBS/iterations'])
    bb, pp = np.nonzero(t.vrow >= 0)
    assert np.abs(V[t.vrow[bb, pp]].T - g[name + '/FBS/V'][:, bb, pp]).max() <= tol
    ll, qq = np.nonzero(t.irow >= 0)
    if i_rows is not None:
        keep = np.array([int(t.irow[l, q]) in i_rows for l, q in zip(ll, qq)])
        ll, qq = ll[keep], qq[keep]
    assert np.abs(I[t.irow[ll, qq]].T - g[name + '/FBS/I'][:, ll, qq]).max() <= tol


def test_the_random_set_covers_what_it_claims():
    """Reverse-declared two-phase lines, every line kind, multi-phase loads, regulators, a transformer, and a bus
    whose phases are rewritten by several children all occur."""
    n_rev = n_sw = n_2ph_load = n_3ph_load = n_reg = n_xfm = n_multi_child = 0
    for name in RANDOM:
        m = rmodel(name)
        for ln in m.lines:
            ph = [int(x) for x in ln.bus1.split('.')[1:]]
            n_rev += ph != sorted(ph)
            n_sw += bool(ln.is_switch)
        n_2ph_load += sum(ld.nphases == 2 for ld in m.loads)
        n_3ph_load += sum(ld.nphases == 3 for ld in m.loads)
        n_reg += len(m.regcontrols)
        n_xfm += sum(t.nphases == 3 for t in m.transformers)
        parents = [ln.bus1.split('.')[0] for ln in m.lines]
        n_multi_child += sum(parents.count(b) >= 3 for b in set(parents))
        assert bool(gold()[name + '/FBS/converged'].all())
    assert n_rev >= 3 and n_sw >= 1 and n_2ph_load >= 5 and n_3ph_load >= 5 and n_reg >= 10 and n_xfm >= 2
    assert n_multi_child >= 5


@pytest.mark.parametrize('name', RANDOM)
def test_fbs_oracle_matches_the_reference(name):
    from oracle.gp_oracle import FBSOracle
    g = gold()
    o = FBSOracle(rmodel(name), ZIPS['test_zip'])
    mult = g[name + '/mult']
    r = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    assert list(r['iterations']) == list(g[name + '/FBS/iterations'])
    for p in ('V', 'I', 'sV', 'Stx', 'Srx'):
        assert np.abs(r[p] - g[f'{name}/FBS/{p}']).max() <= TOL, p
    assert list(o.topo_order) == list(g[name + "/FBS/topo_order"])


@pytest.mark.parametrize('name', [n for n in RANDOM if int(n[3:]) in (0, 1, 2, 3, 4, 5, 6, 7, 12, 13, 14)])
def test_nr3_oracles_match_the_reference(name):
    from oracle.gp_oracle import NR3Oracle
    from oracle.nr3_tree import NR3TreeOracle
    g = gold()
    model = rmodel(name)
    X = g[name + '/NR3/XNR']
    mult = g[name + '/mult'][:len(X)]
    for cls in (NR3Oracle, NR3TreeOracle):
        o = cls(model, ZIPS['test_zip'])
        r = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
        assert list(r['iterations']) == list(g[name + '/NR3/iterations']), cls.__name__
        assert np.abs(r['XNR'] - X).max() <= 1e-11, cls.__name__


@pytest.mark.parametrize('order', ['dfs', 'sweeps'])
@pytest.mark.parametrize('name', RANDOM)
def test_streaming_kernel_loop_over_the_library_schedule_matches_the_reference(name, order):
    from oracle.gp_oracle import FBSOracle
    from test_fbs_order_host import _schedule, emulate
    model, t = _tables(name)
    events, kind = _schedule(t)
    o = FBSOracle(model, ZIPS['test_zip'])
    V, I, its = emulate(t, _packed(t, o, gold()[name + '/mult']), events, kind, order)
    _check_rows(t, V, I, its, name)


@pytest.mark.parametrize('name', RANDOM)
def test_tree_walk_loop_over_the_library_events_matches_the_reference(name):
    from oracle.gp_oracle import FBSOracle
    from test_fbs_walk_host import _events, emulate_walk
    model, t = _tables(name)
    ev = _events(t)
    if len(ev) == 0:
        pytest.skip('feeder has V rows below a regulator that no forward step writes: two-sweep kernel')
    o = FBSOracle(model, ZIPS['test_zip'])
    V, I, its = emulate_walk(t, _packed(t, o, gold()[name + '/mult']), ev)
    stored = {int(e[2 + q]) for e in ev for q in range(3) if (e[1] >> (6 + q)) & 1}
    _check_rows(t, V, I, its, name, i_rows=stored)


@pytest.mark.parametrize('mode', [0, 1, 4])
@pytest.mark.parametrize('name', HOST_COMPILED)
def test_generated_sweeps_compiled_for_the_host_match_the_reference(name, mode, tmp_path):
    """mode: gp_fbs_jit_source's - 0 two sweeps with the I slots in shared memory, 1 I rows in global memory,
    4 the depth-first build."""
    from gigpower_b200 import _cabi
    from oracle.gp_oracle import FBSOracle
    from test_jit_source_on_host import MAIN_DFS, host_program
    model, t = _tables(name)
    d, keep = _cabi.make_fbs_desc(t)
    n = _cabi.gp_fbs_jit_source(C.byref(d), 2, mode, None, 0)
    assert n > 0
    buf = C.create_string_buffer(n + 1)
    assert _cabi.gp_fbs_jit_source(C.byref(d), 2, mode, buf, n + 1) == n
    prog = host_program(buf.value.decode())
    if mode == 4:
        prog = prog[:prog.index('int main(')] + MAIN_DFS
    cpp, exe = tmp_path / 'sweep.cpp', tmp_path / 'sweep'
    cpp.write_text(prog)
    subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-w', '-o', str(exe), str(cpp)], check=True)
    o = FBSOracle(model, ZIPS['test_zip'])
    rows = _packed(t, o, gold()[name + '/mult'])
    S = rows.shape[1]
    fin, fout = tmp_path / 'in.bin', tmp_path / 'out.bin'
    rows[:t.n_srows].astype(np.complex128).tofile(fin)
    subprocess.run([str(exe), str(S), repr(float(t.tolerance)), '100', str(fin), str(fout)], check=True)
    raw = np.fromfile(fout, dtype=np.float64)
    nv, ni = t.n_vrows, max(t.n_irows, 1)
    Vr = raw[:2 * nv * S].view(np.complex128).reshape(nv, S)
    Ir = raw[2 * nv * S:2 * (nv + ni) * S].view(np.complex128).reshape(ni, S)
    its = raw[2 * (nv + ni) * S:].astype(int)
    _check_rows(t, Vr, Ir, its, name, tol=1e-9)


# ---------------------------------------------------------------------------------------------
# per-scenario taps, warm starts and reactive injections on the random feeders:
# tests/golden/ref_random_extras.npz, each case as the reference defines it (oracle/make_golden_random.py: extras)
# ---------------------------------------------------------------------------------------------
TAP_CASES = ['rnd03', 'rnd06', 'rnd08', 'rnd11', 'rnd13']
WARM_CASES = ['rnd02', 'rnd03', 'rnd05', 'rnd10']
WPU_CASES = ['rnd04', 'rnd13', 'rnd09']


def extras():
    if 'x' not in _cache:
        _cache['x'] = np.load(os.path.join(GOLDEN, 'ref_random_extras.npz'))
    return _cache['x']


def test_the_extras_set_holds_a_scenario_the_reference_does_not_converge():
    """Taps all at -16 on the 40-bus feeder: the reference stops at maxiter = 100 (a slow geometric approach that
    would take 159 sweeps), so the maxiter exit is pinned by the reference too."""
    it = extras()['rnd11/taps/FBS/iterations']
    assert it.max() == 100 and (it < 100).sum() == 4 and np.isfinite(extras()['rnd11/taps/FBS/V']).all()


@pytest.mark.parametrize('name', TAP_CASES)
def test_oracles_with_per_scenario_taps_match_reference_objects(name):
    from oracle.gp_oracle import FBSOracle, NR3Oracle, OracleCircuit, solve_with_taps
    from oracle.nr3_tree import NR3TreeOracle
    x = extras()
    model = rmodel(name)
    assert [rc.name for rc in model.regcontrols] == list(x[name + '/taps/regs'])
    c = OracleCircuit(model)
    mult, taps = x[name + '/taps/mult'], x[name + '/taps/taps']
    spu = c.spu_matrix(c.load_kw * mult, c.load_kvar * mult)
    r = solve_with_taps(FBSOracle, model, spu, taps, ZIPS['test_zip'])
    assert list(r['iterations']) == list(x[name + '/taps/FBS/iterations'])
    assert np.abs(r['V'] - x[name + '/taps/FBS/V']).max() <= TOL
    assert np.abs(r['I'] - x[name + '/taps/FBS/I']).max() <= TOL
    if name + '/taps/NR3/XNR' in x.files:
        X = x[name + '/taps/NR3/XNR']
        for cls, tol in ((NR3Oracle, 1e-11), (NR3TreeOracle, 1e-10)):
            n = solve_with_taps(cls, model, spu[:len(X)], taps[:len(X)], ZIPS['test_zip'])
            assert list(n['iterations']) == list(x[name + '/taps/NR3/iterations'])
            assert np.abs(n['XNR'] - X).max() <= tol


@pytest.mark.parametrize('name', WARM_CASES)
def test_oracles_warm_started_match_a_reused_reference_object(name):
    from oracle.gp_oracle import FBSOracle, NR3Oracle
    x = extras()
    model = rmodel(name)
    o = FBSOracle(model, ZIPS['test_zip'])
    c = o.c
    mult = x[name + '/warm/mult']
    V0 = I0 = None
    for h in range(len(mult)):
        r = o.solve(c.spu_matrix(c.load_kw * mult[h], c.load_kvar * mult[h])[None], V0=V0, I0=I0)
        assert int(r['iterations'][0]) == int(x[name + '/warm/FBS/iterations'][h]), h
        assert np.abs(r['V'][0] - x[name + '/warm/FBS/V'][h]).max() <= TOL
        assert np.abs(r['I'][0] - x[name + '/warm/FBS/I'][h]).max() <= TOL
        V0, I0 = r['V'], r['I']
    if name + '/warm/NR3/XNR' in x.files:
        n = NR3Oracle(model, ZIPS['test_zip'])
        X0 = None
        for h in range(len(x[name + '/warm/NR3/iterations'])):
            r = n.solve(c.spu_matrix(c.load_kw * mult[h], c.load_kvar * mult[h])[None], X0=X0)
            assert int(r['iterations'][0]) == int(x[name + '/warm/NR3/iterations'][h]), h
            assert np.abs(r['XNR'][0] - x[name + '/warm/NR3/XNR'][h]).max() <= 1e-11
            X0 = r['XNR']


@pytest.mark.parametrize('name', WPU_CASES)
def test_fbs_oracle_with_reactive_injection_matches_reference_objects(name):
    from oracle.gp_oracle import FBSOracle
    x = extras()
    o = FBSOracle(rmodel(name), ZIPS['test_zip'])
    c = o.c
    mult = x[name + '/wpu/mult']
    r = o.solve(c.spu_matrix(c.load_kw * mult, c.load_kvar * mult), wpu=x[name + '/wpu/wpu'])
    assert list(r['iterations']) == list(x[name + '/wpu/FBS/iterations'])
    for p in ('V', 'I', 'sV'):
        assert np.abs(r[p] - x[f'{name}/wpu/FBS/{p}']).max() <= TOL, p


@pytest.mark.parametrize('name', TAP_CASES)
def test_generated_sweep_with_per_scenario_taps_matches_reference_objects(name, tmp_path):
    """The EX build of the specialised kernel (gp_fbs_jit_source mode 2), compiled for the host, with the golden's
    taps as per-scenario ratio rows."""
    from gigpower_b200 import _cabi
    from oracle.gp_oracle import FBSOracle
    from test_jit_source_on_host import MAIN_EX, host_program
    x = extras()
    model, t = _tables(name)
    d, keep = _cabi.make_fbs_desc(t)
    n = _cabi.gp_fbs_jit_source(C.byref(d), 2, 2, None, 0)
    buf = C.create_string_buffer(n + 1)
    assert _cabi.gp_fbs_jit_source(C.byref(d), 2, 2, buf, n + 1) == n
    prog = host_program(buf.value.decode())
    prog = prog[:prog.index('int main(')] + MAIN_EX
    cpp, exe = tmp_path / 'sweep_ex.cpp', tmp_path / 'sweep_ex'
    cpp.write_text(prog)
    subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-w', '-o', str(exe), str(cpp)], check=True)
    o = FBSOracle(model, ZIPS['test_zip'])
    taps = x[name + '/taps/taps']
    rows = _packed(t, o, x[name + '/taps/mult'])
    S = rows.shape[1]
    g = (taps.astype(float) + 16) / 32
    g = g * (1.10 - .90)
    g = g + .90                                             # voltage_regulator.py:23-39, same operation order
    gam = np.ascontiguousarray(g.T[np.asarray(t.grow_vr, dtype=int)])       # [n_gamma_rows, S]
    fin, fg, fout = tmp_path / 'in.bin', tmp_path / 'gam.bin', tmp_path / 'out.bin'
    rows[:t.n_srows].astype(np.complex128).tofile(fin)
    gam.astype(np.float64).tofile(fg)
    subprocess.run([str(exe), str(S), repr(float(t.tolerance)), '100', str(fin), str(fg), str(fout), str(gam.shape[0])], check=True)
    raw = np.fromfile(fout, dtype=np.float64)
    nv, ni = t.n_vrows, max(t.n_irows, 1)
    Vr = raw[:2 * nv * S].view(np.complex128).reshape(nv, S)
    its = raw[2 * (nv + ni) * S:].astype(int)
    want_it = x[name + '/taps/FBS/iterations']
    assert list(its) == list(want_it)
    bb, pp = np.nonzero(t.vrow >= 0)
    assert np.abs(Vr[t.vrow[bb, pp]].T - x[name + '/taps/FBS/V'][:, bb, pp]).max() <= 1e-9


DROOP_CASES = ['rnd09', 'rnd13', 'rnd10', 'rnd11']
WPU_NR3_CASES = ['rnd04', 'rnd13']


@pytest.mark.parametrize('name', DROOP_CASES)
def test_fbs_oracle_with_droop_inside_the_iteration_matches_the_reference_controllers(name):
    """The reference's SolutionFBS with its own VoltVARController objects consulted before every calc_sV
    (oracle/make_golden_vvc_iter.py has the construction), on random feeders; the injections saturate on two of them."""
    from oracle.gp_oracle import FBSOracle, vvc_get_Q
    x = extras()
    o = FBSOracle(rmodel(name), ZIPS['test_zip'])
    mult = x[name + '/droop/mult']
    ctl = [tuple(int(v) for v in row) for row in x[name + '/droop/ctl']]
    pts, (minQ, maxQ) = tuple(x['droop/points']), x['droop/q']
    r = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult), droop=(ctl, pts, minQ, maxQ))
    assert list(r['iterations']) == list(x[name + '/droop/iterations'])
    for k in ('V', 'I', 'sV'):
        assert np.abs(r[k] - x[f'{name}/droop/{k}']).max() <= TOL, k
    Q = np.stack([vvc_get_Q(np.abs(r['V'][:, k, p]), pts, minQ, maxQ) for k, p in ctl], axis=1)
    assert np.abs(Q - x[name + '/droop/Q']).max() <= TOL and np.abs(Q).max() > 1e-2


@pytest.mark.parametrize('name', WPU_NR3_CASES)
def test_nr3_tree_oracle_with_wpu_matches_the_reference_on_random_feeders(name):
    """Reference SolutionNR3 objects whose circuit reports a non-zero wpu, after the reference's own
    change_KCL_matrices has written it into their KCL constants (oracle/make_golden_wpu_nr3.py)."""
    from oracle.nr3_tree import NR3TreeOracle
    x = extras()
    o = NR3TreeOracle(rmodel(name), ZIPS['test_zip'])
    mult, wpu = x[name + '/wpu_nr3/mult'], x[name + '/wpu_nr3/wpu']
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    r = o.solve(spu, wpu=wpu)
    assert list(r['iterations']) == list(x[name + '/wpu_nr3/iterations'])
    assert np.abs(r['XNR'] - x[name + '/wpu_nr3/XNR']).max() <= 1e-11
    assert np.abs(o.solve(spu)['XNR'] - x[name + '/wpu_nr3/XNR']).max() >= 1e-5


# ---------------------------------------------------------------------------------------------
# a feeder on which the reference itself has two answers (oracle/make_golden_two_banks.py)
# ---------------------------------------------------------------------------------------------
def test_two_regulator_banks_off_one_bus_follow_the_order_of_definition(tmp_path):
    """Bus b1 feeds two regulator banks.  The reference de-duplicates a bus's regulator edges with ``list(set(...))``
    (voltage_regulator_group.py:60-72), so the banks' order in its adjacency list - and with it the topo order, the
    child that rewrites b1 last and the voltages of the source and of b1 (3.6e-7 apart) - follows PYTHONHASHSEED: the
    golden file holds both answers, each from several hash seeds.  The flattener, the oracle, the kernel loops over the
    library's schedules and the generated sweeps all give the one with the banks in the order of definition."""
    from gigpower_b200 import _cabi
    from oracle.gp_oracle import FBSOracle
    from test_fbs_order_host import _schedule, emulate
    from test_jit_source_on_host import MAIN_DFS, host_program
    g = np.load(os.path.join(GOLDEN, 'ref_two_banks.npz'))
    assert int(g['n_variants']) == 2
    ours, other = ('v0', 'v1') if bool(g['v0/definition_order']) else ('v1', 'v0')
    assert bool(g[ours + '/definition_order']) and not bool(g[other + '/definition_order'])
    assert len(g[ours + '/hash_seeds']) >= 2 and len(g[other + '/hash_seeds']) >= 2
    assert np.abs(g[ours + '/V'] - g[other + '/V']).max() > 1e-7            # the two answers really differ
    model, t = _tables('two_banks')
    o = FBSOracle(model, ZIPS['test_zip'])
    assert list(o.topo_order) == list(g[ours + '/topo_order'])
    assert [model.bus_names.index(n) for n in t.topo_order] == list(g[ours + '/topo_order'])
    r = o.solve(o.c.spu_matrix(o.c.load_kw[None], o.c.load_kvar[None]))
    assert int(r['iterations'][0]) == int(g[ours + '/iterations'])
    assert np.abs(r['V'][0] - g[ours + '/V']).max() <= TOL and np.abs(r['I'][0] - g[ours + '/I']).max() <= TOL
    rows = _packed(t, o, np.ones((1, len(o.c.load_names))))
    bb, pp = np.nonzero(t.vrow >= 0)
    events, kind = _schedule(t)
    for order in ('dfs', 'sweeps'):
        V, I, its = emulate(t, rows, events, kind, order)
        assert int(its[0]) == int(g[ours + '/iterations'])
        assert np.abs(V[t.vrow[bb, pp], 0] - g[ours + '/V'][bb, pp]).max() <= TOL
    d, keep = _cabi.make_fbs_desc(t)
    for mode in (0, 4):
        n = _cabi.gp_fbs_jit_source(C.byref(d), 2, mode, None, 0)
        buf = C.create_string_buffer(n + 1)
        assert _cabi.gp_fbs_jit_source(C.byref(d), 2, mode, buf, n + 1) == n
        prog = host_program(buf.value.decode())
        if mode == 4:
            prog = prog[:prog.index('int main(')] + MAIN_DFS
        cpp, exe = tmp_path / f'sweep{mode}.cpp', tmp_path / f'sweep{mode}'
        cpp.write_text(prog)
        subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-w', '-o', str(exe), str(cpp)], check=True)
        fin, fout = tmp_path / 'in.bin', tmp_path / 'out.bin'
        rows[:t.n_srows].astype(np.complex128).tofile(fin)
        subprocess.run([str(exe), '1', repr(float(t.tolerance)), '100', str(fin), str(fout)], check=True)
        raw = np.fromfile(fout, dtype=np.float64)
        Vr = raw[:2 * t.n_vrows].view(np.complex128).reshape(t.n_vrows, 1)
        assert int(raw[2 * (t.n_vrows + max(t.n_irows, 1)):][0]) == int(g[ours + '/iterations'])
        assert np.abs(Vr[t.vrow[bb, pp], 0] - g[ours + '/V'][bb, pp]).max() <= 1e-9
        assert np.abs(Vr[t.vrow[bb, pp], 0] - g[other + '/V'][bb, pp]).max() > 1e-7


# ---------------------------------------------------------------------------------------------
# the PRODUCT's NR3 tables (flatten_nr3) exercised without a GPU: tests/nr3_tables_oracle.py
# ---------------------------------------------------------------------------------------------
def _nr3_tables(model):
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.flatten import flatten_nr3
    from gigpower_b200.solution import Solution
    Solution.set_zip_values(np.asarray(ZIPS['test_zip']))
    return flatten_nr3(Circuit(model))


def _rows_nr3(t, spu):
    b, p = np.nonzero(t.srow >= 0)
    rows = np.zeros((max(t.n_srows, 1), spu.shape[0]), dtype=complex)
    rows[t.srow[b, p]] = spu[:, b, p].T
    return rows


@pytest.mark.parametrize('name', [n for n in RANDOM if int(n[3:]) in (0, 1, 2, 3, 4, 5, 6, 7, 12, 13, 14)])
def test_newton_over_the_products_nr3_tables_matches_the_reference_on_random_feeders(name):
    """The tree elimination run over flatten_nr3's tables (topology, blocks, ratios, capacitors, load rows, XNR map)
    gives the reference's X and iteration counts."""
    from nr3_tables_oracle import NR3TablesOracle
    from oracle.gp_oracle import OracleCircuit
    g = gold()
    model = rmodel(name)
    t = _nr3_tables(model)
    X = g[name + '/NR3/XNR']
    mult = g[name + '/mult'][:len(X)]
    c = OracleCircuit(model, ZIPS['test_zip'])
    r = NR3TablesOracle(t).solve_rows(_rows_nr3(t, c.spu_matrix(c.load_kw * mult, c.load_kvar * mult)))
    assert list(r['iterations']) == list(g[name + '/NR3/iterations'])
    assert np.abs(r['XNR'] - X).max() <= 1e-11


@pytest.mark.parametrize('feeder', ['IEEE_13_Bus_allwye', 'IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_34_Bus_allwye',
                                    'IEEE_34_Bus_allwye_noxfm_noreg', 'IEEE_37_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg'])
def test_newton_over_the_products_nr3_tables_matches_the_reference_snapshots(feeder):
    from helpers import load_model, snapshots
    from nr3_tables_oracle import NR3TablesOracle
    from oracle.gp_oracle import OracleCircuit
    model = load_model(feeder)
    t = _nr3_tables(model)
    c = OracleCircuit(model, ZIPS['test_zip'])
    r = NR3TablesOracle(t).solve_rows(_rows_nr3(t, c.spu_matrix()[None]))
    k = f'{feeder}/test_zip/NR3/'
    assert int(r['iterations'][0]) == int(snapshots()[k + 'iterations'])
    assert np.abs(r['XNR'][0] - snapshots()[k + 'XNR'][:, 0]).max() <= 1e-11


@pytest.mark.parametrize('source,name', [('taps', 'IEEE_13_Bus_allwye'), ('taps', 'IEEE_37_Bus_allwye'),
                                         ('random', 'rnd03'), ('random', 'rnd06'), ('random', 'rnd13')])
def test_per_scenario_gamma_rows_of_the_nr3_tables_match_reference_objects(source, name):
    """``grow`` (step, phase -> row of the per-scenario ratio array) and ``grow_vr`` (row -> regulator) of
    flatten_nr3, as ``Solution.gamma_rows_from_taps`` uses them: the Newton step over the product's tables with every
    regulated phase's ratio taken from its scenario's gamma row meets one reference object per scenario."""
    from helpers import load_model, taps_golden
    from nr3_tables_oracle import NR3TablesOracle
    from oracle.gp_oracle import OracleCircuit
    if source == 'taps':
        g, model, key = taps_golden(), load_model(name), name
        X, its, mult, taps = g[key + '/NR3/XNR'], g[key + '/NR3/iterations'], g[key + '/mult'], g[key + '/taps']
    else:
        g, model, key = extras(), rmodel(name), name + '/taps'
        X, its, mult, taps = g[key + '/NR3/XNR'], g[key + '/NR3/iterations'], g[key + '/mult'], g[key + '/taps']
    t = _nr3_tables(model)
    c = OracleCircuit(model, ZIPS['test_zip'])
    n = len(X)
    rows = _rows_nr3(t, c.spu_matrix(c.load_kw * mult[:n], c.load_kvar * mult[:n]))
    gq = (taps[:n].astype(float) + 16) / 32
    gq = gq * (1.10 - .90)
    gq = gq + .90                                           # voltage_regulator.py:23-39, same operation order
    gam = gq.T[np.asarray(t.grow_vr, dtype=int)]            # [n_gamma_rows, S], as gamma_rows_from_taps builds it
    assert gam.shape[0] == int((t.grow >= 0).sum()) > 0
    for s in range(n):
        o = NR3TablesOracle(t)
        for i, e in enumerate(o.edges):
            for p in range(3):
                if t.grow[i, p] >= 0:
                    assert e['kind'] == 'vr' and e['pm'][p]
                    e['gamma'][p] = gam[t.grow[i, p], s]
        r = o.solve_rows(rows[:, s:s + 1])
        assert int(r['iterations'][0]) == int(its[s]), s
        assert np.abs(r['XNR'][0] - X[s]).max() <= 1e-10, s

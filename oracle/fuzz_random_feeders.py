"""Fuzz: many seeded random radial feeders through the UNMODIFIED reference and through everything that can run
without a GPU -- TEST INFRASTRUCTURE, build container only (needs ``/root/reference``).

For every seed: ``make_golden_random.make_feeder`` writes a feeder, the reference's own ``SolutionFBS`` (and, on small
feeders, ``SolutionNR3``) solves it at nominal load and at one random load vector, and the result is compared with

* the FBS oracle and the flattener's topo order (1e-12, equal iteration counts);
* the streaming kernel's loop run in NumPy over ``gp_fbs_schedule`` in the depth-first order, and the tree-walk
  kernel's loop over ``gp_fbs_walk_events`` (tests/test_fbs_order_host.py, tests/test_fbs_walk_host.py);
* the dense and the tree NR3 oracles (1e-11), and the tree elimination run over the product's own NR3 tables
  (``flatten_nr3``; tests/nr3_tables_oracle.py).

Sixteen of these feeders are committed as fixtures (tests/golden/random_feeders, tests/test_random_feeders*.py); this
script is the wider net.  Usage: python oracle/fuzz_random_feeders.py [--jit] [--nr3] [--big] [--zip] [first_seed] [count]   (--jit: also the
generated sweeps of the feeder-specialised kernel, compiled for the host with g++; prints one summary line
per 25 feeders and every mismatch; exit status 1 if there was one)
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                      # noqa: E402
from make_golden_random import make_feeder                    # noqa: E402

sys.path.insert(0, os.path.join(mg.REPO, 'tests'))
from test_fbs_order_host import _schedule, emulate            # noqa: E402
from test_fbs_walk_host import _events, emulate_walk          # noqa: E402
from gigpower_b200.circuit import Circuit                     # noqa: E402
from gigpower_b200.flatten import flatten_fbs                 # noqa: E402
from gigpower_b200.solution import Solution                   # noqa: E402
from oracle.gp_oracle import FBSOracle, NR3Oracle             # noqa: E402
from oracle.nr3_tree import NR3TreeOracle                     # noqa: E402


def differs(a, b, tol):
    """max |a - b| / max(1, |b|) > tol over the finite entries, or a different pattern of non-finite ones."""
    a, b = np.asarray(a), np.asarray(b)
    fa, fb = np.isfinite(a), np.isfinite(b)
    if not np.array_equal(fa, fb):
        return True
    # relative above 1 (rel_err of tests/helpers.py): a diverging sweep reaches |V| of 1e60 with heavy constant-impedance
    # loads (seeds 80109, 81081) and still agrees to 1e-15 of that
    return bool(fa.any()) and float((np.abs(a[fa] - b[fa]) / np.maximum(1.0, np.abs(b[fa]))).max()) > tol


def one(seed, tmp, stats):
    rng = np.random.Generator(np.random.PCG64(seed * 7 + 1))
    nb = int(rng.integers(45, 160)) if BIG else int(rng.integers(4, 45))
    reg, xfm = bool(rng.random() < 0.35), bool(rng.random() < 0.25 and nb >= 8)
    name = f'fz{seed}'
    text, taps = make_feeder(name, seed, nb, reg, xfm, exotic=bool(seed % 2))
    fp = os.path.join(tmp, name + '.dss')
    with open(fp, 'w') as fh:
        fh.write(text)
    z = mg.ZIPS['test_zip']
    if ZIPMIX:                                   # --zip: one of the reference's three ZIP sets or a random split
        u = rng.random(3)
        u = (u / u.sum()).round(3)
        u[2] = 1.0 - u[0] - u[1]
        z = [mg.ZIPS['constant_impedance'], mg.ZIPS['constant_power'], mg.ZIPS['test_zip'],
             [float(u[0]), float(u[1]), float(u[2])] * 2 + [0.8]][int(rng.integers(4))]
    mg.dss.TAPS = taps
    model = mg.read_dss(fp, taps)
    mult = np.vstack([np.ones(len(model.loads)), rng.uniform(0.3, 1.6, size=len(model.loads))])
    ref = [mg.ref_fbs(fp, z, m) for m in mult]
    Vr, Ir = np.array([s.V for s in ref]), np.array([s.I for s in ref])
    itr = [s.iterations for s in ref]
    bad = []
    # a sweep the reference leaves unconverged at maxiter is either a slow geometric approach (reproduced to 1e-12) or
    # a voltage-collapse orbit with |V| of several pu, on which rounding differences of 1e-16 grow to 1e-9 in 100
    # sweeps (seeds 31275, 31371, 31395, 31413) and to 2e-6 on seed 41207: those feeders are held to 1e-3, i.e. to
    # the iteration count and to being on the same orbit
    tol = 1e-12 if all(s.converged for s in ref) else 1e-3
    o = FBSOracle(model, z)
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    r = o.solve(spu)
    if list(r['iterations']) != itr or differs(r['V'], Vr, tol) or differs(r['I'], Ir, tol):
        bad.append('fbs oracle')
    if list(o.topo_order) != [ref[0].circuit.buses.get_idx(n) for n in ref[0].topo_order]:
        bad.append('topo order')
    Solution.set_zip_values(np.asarray(z))
    ours = Circuit(model)
    # the Circuit accessors the drop-in mirrors (circuit.py:90-205, circuit_element_group.py:39-91), on the nominal object
    rc = mg.SolutionFBS(fp, zip_v=np.asarray(z)).circuit
    try:
        for grp in ('buses', 'lines', 'loads', 'capacitors', 'transformers', 'voltage_regulators'):
            a, b_ = getattr(ours, grp), getattr(rc, grp)
            assert a.num_elements == b_.num_elements and [str(x) for x in a.all_names()] == [str(x) for x in b_.all_names()], grp
        assert np.array_equal(ours.buses.get_phase_matrix('rows'), rc.buses.get_phase_matrix('rows'))
        for m in ('get_spu_matrix', 'get_cappu_matrix', 'get_aPQ_matrix', 'get_aI_matrix', 'get_aZ_matrix', 'get_wpu_matrix'):
            assert np.abs(np.asarray(getattr(ours, m)()) - np.asarray(getattr(rc, m)())).max() <= 1e-15, m
        assert list(ours.get_tx_idx_matrix()) == list(rc.get_tx_idx_matrix()) and list(ours.get_rx_idx_matrix()) == list(rc.get_rx_idx_matrix())
        assert ours.get_total_lines() == rc.get_total_lines()
        assert np.abs(ours.get_nominal_bus_powers().to_numpy() - rc.get_nominal_bus_powers().to_numpy()).max() <= 1e-15
        for la, lb in zip(ours.lines.get_elements(), rc.lines.get_elements()):
            assert np.abs(la.FZpu - lb.FZpu).max() <= 1e-15 * max(1.0, np.abs(lb.FZpu).max()), 'FZpu'
    except AssertionError as e:
        bad.append(f'Circuit accessors: {e}')
    t = flatten_fbs(ours)
    b, p = np.nonzero(t.srow >= 0)
    rows = np.zeros((max(t.n_srows, 1), len(mult)), dtype=complex)
    rows[t.srow[b, p]] = spu[:, b, p].T
    bb, pp = np.nonzero(t.vrow >= 0)
    ll, qq = np.nonzero(t.irow >= 0)
    events, kind = _schedule(t)
    V, I, its = emulate(t, rows, events, kind, 'dfs')
    if list(its) != itr or differs(V[t.vrow[bb, pp]].T, Vr[:, bb, pp], tol) \
            or differs(I[t.irow[ll, qq]].T, Ir[:, ll, qq], tol):
        bad.append('kernel loop over gp_fbs_schedule')
    ev = _events(t)
    if len(ev):
        V, I, its = emulate_walk(t, rows, ev)
        if list(its) != itr or differs(V[t.vrow[bb, pp]].T, Vr[:, bb, pp], tol):
            bad.append('walk loop over gp_fbs_walk_events')
        stats['walked'] += 1
    if nb <= NR3_MAX_BUSES:
        n, itn = mg.ref_nr3(fp, z, mult[1])
        X = n.XNR[:, 0].copy()
        del n
        for cls in (NR3Oracle, NR3TreeOracle):
            q = cls(model, z)
            rq = q.solve(q.c.spu_matrix(q.c.load_kw * mult[1:], q.c.load_kvar * mult[1:]))
            # (a Newton run the reference itself leaves unconverged at 100 iterations wanders chaotically through
            # near-singular Jacobians - |X| in the tens, seeds 21243 and 30057 -: rounding differences are amplified
            # without bound there and nothing is compared: not the values, and not the count either, because the loop
            # stops at the first residual that holds a NaN and WHEN such an orbit overflows is as chaotic as the orbit
            # - the reference at iteration 83 with seed 31191, the dense oracle at 87 with seed 32067)
            if np.isfinite(X).all() and itn < 100 and (int(rq['iterations'][0]) != itn or differs(rq['XNR'][0], X, 1e-11)):
                bad.append(cls.__name__)
        if np.isfinite(X).all() and itn < 100:
            # the same Newton step run over the PRODUCT's NR3 tables (tests/nr3_tables_oracle.py)
            from nr3_tables_oracle import NR3TablesOracle
            from gigpower_b200.flatten import flatten_nr3
            tn = flatten_nr3(ours)
            bn, pn = np.nonzero(tn.srow >= 0)
            rn = np.zeros((max(tn.n_srows, 1), 1), dtype=complex)
            rn[tn.srow[bn, pn]] = spu[1:, bn, pn].T
            rq = NR3TablesOracle(tn).solve_rows(rn)
            if int(rq['iterations'][0]) != itn or differs(rq['XNR'][0], X, 1e-11):
                bad.append('Newton over flatten_nr3 tables')
        stats['nr3'] += 1
        stats['nr3_unconverged'] = stats.get('nr3_unconverged', 0) + int(itn >= 100)
    if JIT:
        # the GENERATED sweeps of the feeder-specialised kernel (two sweeps with the I slots in shared / global memory,
        # the depth-first build), compiled for the host as tests/test_jit_source_on_host.py does
        import ctypes as C
        import subprocess
        from gigpower_b200 import _cabi
        from test_jit_source_on_host import MAIN_DFS, host_program
        d, keep = _cabi.make_fbs_desc(t)
        for mode in (0, 1, 4):
            n = _cabi.gp_fbs_jit_source(C.byref(d), 2, mode, None, 0)
            if n <= 0:
                stats['jit_refused'] = stats.get('jit_refused', 0) + 1      # depth-first: more than 64 steps
                continue
            buf = C.create_string_buffer(n + 1)
            _cabi.gp_fbs_jit_source(C.byref(d), 2, mode, buf, n + 1)
            prog = host_program(buf.value.decode())
            if mode == 4:
                prog = prog[:prog.index('int main(')] + MAIN_DFS
            cpp, exe = os.path.join(tmp, f'{name}_{mode}.cpp'), os.path.join(tmp, f'{name}_{mode}')
            with open(cpp, 'w') as fh:
                fh.write(prog)
            subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-w', '-o', exe, cpp], check=True)
            fin, fout = os.path.join(tmp, 'in.bin'), os.path.join(tmp, 'out.bin')
            rows[:t.n_srows].astype(np.complex128).tofile(fin)
            subprocess.run([exe, str(len(mult)), repr(float(t.tolerance)), '100', fin, fout], check=True)
            raw = np.fromfile(fout, dtype=np.float64)
            nv, ni, S = t.n_vrows, max(t.n_irows, 1), len(mult)
            Vj = raw[:2 * nv * S].view(np.complex128).reshape(nv, S)
            Ij = raw[2 * nv * S:2 * (nv + ni) * S].view(np.complex128).reshape(ni, S)
            itj = raw[2 * (nv + ni) * S:].astype(int)
            tj = max(tol, 1e-9)                  # the fast path's rsqrt / FMA arithmetic: the 1e-9 bar of the GPU tests
            if list(itj) != itr or differs(Vj[t.vrow[bb, pp]].T, Vr[:, bb, pp], tj) or differs(Ij[t.irow[ll, qq]].T, Ir[:, ll, qq], tj):
                bad.append(f'generated sweep, mode {mode}')
            os.remove(exe), os.remove(cpp)
        if taps:
            # the EX build (mode 2) with per-scenario taps against one reference object per scenario built with them
            from test_jit_source_on_host import MAIN_EX
            regs = [rc.name for rc in model.regcontrols]
            tp = rng.integers(-16, 17, size=(len(mult), len(regs)))
            refs = []
            for i in range(len(mult)):
                mg.dss.TAPS = {r: int(v) for r, v in zip(regs, tp[i])}
                refs.append(mg.ref_fbs(fp, z, mult[i]))
            mg.dss.TAPS = taps
            n = _cabi.gp_fbs_jit_source(C.byref(d), 2, 2, None, 0)
            buf = C.create_string_buffer(n + 1)
            _cabi.gp_fbs_jit_source(C.byref(d), 2, 2, buf, n + 1)
            prog = host_program(buf.value.decode())
            prog = prog[:prog.index('int main(')] + MAIN_EX
            cpp, exe = os.path.join(tmp, f'{name}_ex.cpp'), os.path.join(tmp, f'{name}_ex')
            with open(cpp, 'w') as fh:
                fh.write(prog)
            subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-w', '-o', exe, cpp], check=True)
            gq = (tp.astype(float) + 16) / 32
            gq = gq * (1.10 - .90)
            gq = gq + .90                                   # voltage_regulator.py:23-39, same operation order
            gam = np.ascontiguousarray(gq.T[np.asarray(t.grow_vr, dtype=int)])
            fin, fg, fout = os.path.join(tmp, 'in.bin'), os.path.join(tmp, 'gam.bin'), os.path.join(tmp, 'out.bin')
            rows[:t.n_srows].astype(np.complex128).tofile(fin)
            gam.astype(np.float64).tofile(fg)
            subprocess.run([exe, str(len(mult)), repr(float(t.tolerance)), '100', fin, fg, fout, str(gam.shape[0])], check=True)
            raw = np.fromfile(fout, dtype=np.float64)
            Vj = raw[:2 * nv * S].view(np.complex128).reshape(nv, S)
            itj = raw[2 * (nv + ni) * S:].astype(int)
            Vt = np.array([q.V for q in refs])
            tt = 1e-9 if all(q.converged for q in refs) else 1e-3
            if list(itj) != [q.iterations for q in refs] or differs(Vj[t.vrow[bb, pp]].T, Vt[:, bb, pp], tt):
                bad.append('generated sweep, mode 2 (per-scenario taps)')
            os.remove(exe), os.remove(cpp)
            stats['jit_taps'] = stats.get('jit_taps', 0) + 1
        stats['jit'] = stats.get('jit', 0) + 1
    stats['n'] += 1
    stats['regs'] += bool(taps)
    stats['xfm'] += bool(model.transformers) and any(tr.nphases == 3 for tr in model.transformers)
    stats['unconverged'] += sum(not s.converged for s in ref)
    stats['max_it'] = max(stats['max_it'], max(itr))
    stats['nonfinite'] = stats.get('nonfinite', 0) + int(not np.isfinite(Vr).all())
    if bad:
        print(f'MISMATCH seed {seed} ({nb} buses, taps {taps}): {bad}', flush=True)
        with open(os.path.join(mg.REPO, 'gpurun_out', f'fuzz_{name}.dss'), 'w') as fh:
            fh.write(text)
    return not bad


NR3_MAX_BUSES = 18 if '--nr3' in sys.argv else 14         # --nr3: NR3 on more (bigger) feeders
if '--nr3' in sys.argv:
    sys.argv.remove('--nr3')
ZIPMIX = '--zip' in sys.argv
if ZIPMIX:
    sys.argv.remove('--zip')
BIG = '--big' in sys.argv                                     # --big: 45-160 buses instead of 4-44
if BIG:
    sys.argv.remove('--big')
JIT = '--jit' in sys.argv
if JIT:
    sys.argv.remove('--jit')


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    tmp = tempfile.mkdtemp(prefix='gp_fuzz_')
    os.makedirs(os.path.join(mg.REPO, 'gpurun_out'), exist_ok=True)
    stats = dict(n=0, regs=0, xfm=0, nr3=0, walked=0, unconverged=0, max_it=0)
    n_bad = 0
    for k in range(count):
        try:
            n_bad += not one(first + k, tmp, stats)
        except Exception as e:                                        # a feeder the reference itself cannot run
            print(f'seed {first + k}: {type(e).__name__}: {e}', flush=True)
            stats['errors'] = stats.get('errors', 0) + 1
        if (k + 1) % 25 == 0:
            print(f'{k + 1} feeders, {n_bad} mismatches, {stats}', flush=True)
    print(f'DONE seeds {first}..{first + count - 1}: {n_bad} mismatches, {stats}')
    sys.exit(1 if n_bad else 0)


if __name__ == '__main__':
    main()

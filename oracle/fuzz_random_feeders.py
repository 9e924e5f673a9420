"""Fuzz: many seeded random radial feeders through the UNMODIFIED reference and through everything that can run
without a GPU -- TEST INFRASTRUCTURE, build container only (needs ``/root/reference``).

For every seed: ``make_golden_random.make_feeder`` writes a feeder, the reference's own ``SolutionFBS`` (and, on small
feeders, ``SolutionNR3``) solves it at nominal load and at one random load vector, and the result is compared with

* the FBS oracle and the flattener's topo order (1e-12, equal iteration counts);
* the streaming kernel's loop run in NumPy over ``gp_fbs_schedule`` in the depth-first order, and the tree-walk
  kernel's loop over ``gp_fbs_walk_events`` (tests/test_fbs_order_host.py, tests/test_fbs_walk_host.py);
* the dense and the tree NR3 oracles (1e-11).

Sixteen of these feeders are committed as fixtures (tests/golden/random_feeders, tests/test_random_feeders*.py); this
script is the wider net.  Usage: python oracle/fuzz_random_feeders.py [first_seed] [count]   (prints one summary line
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
    """max |a - b| > tol over the finite entries, or a different pattern of non-finite ones."""
    a, b = np.asarray(a), np.asarray(b)
    fa, fb = np.isfinite(a), np.isfinite(b)
    if not np.array_equal(fa, fb):
        return True
    return bool(fa.any()) and float(np.abs(a[fa] - b[fa]).max()) > tol


def one(seed, tmp, stats):
    rng = np.random.Generator(np.random.PCG64(seed * 7 + 1))
    nb = int(rng.integers(4, 45))
    reg, xfm = bool(rng.random() < 0.35), bool(rng.random() < 0.25 and nb >= 8)
    name = f'fz{seed}'
    text, taps = make_feeder(name, seed, nb, reg, xfm, exotic=bool(seed % 2))
    fp = os.path.join(tmp, name + '.dss')
    with open(fp, 'w') as fh:
        fh.write(text)
    z = mg.ZIPS['test_zip']
    mg.dss.TAPS = taps
    model = mg.read_dss(fp, taps)
    mult = np.vstack([np.ones(len(model.loads)), rng.uniform(0.3, 1.6, size=len(model.loads))])
    ref = [mg.ref_fbs(fp, z, m) for m in mult]
    Vr, Ir = np.array([s.V for s in ref]), np.array([s.I for s in ref])
    itr = [s.iterations for s in ref]
    bad = []
    # a sweep the reference leaves unconverged at maxiter is either a slow geometric approach (reproduced to 1e-12) or
    # a voltage-collapse orbit with |V| of several pu, on which rounding differences of 1e-16 grow to 1e-9 in 100
    # sweeps (seeds 31275, 31371, 31395, 31413): those feeders are held to 1e-6
    tol = 1e-12 if all(s.converged for s in ref) else 1e-6
    o = FBSOracle(model, z)
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    r = o.solve(spu)
    if list(r['iterations']) != itr or differs(r['V'], Vr, tol) or differs(r['I'], Ir, tol):
        bad.append('fbs oracle')
    if list(o.topo_order) != [ref[0].circuit.buses.get_idx(n) for n in ref[0].topo_order]:
        bad.append('topo order')
    Solution.set_zip_values(np.asarray(z))
    t = flatten_fbs(Circuit(model))
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
    if nb <= 14:
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
        stats['nr3'] += 1
        stats['nr3_unconverged'] = stats.get('nr3_unconverged', 0) + int(itn >= 100)
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

"""Golden vectors for the two AUTHORED feeders from the UNMODIFIED reference -- TEST INFRASTRUCTURE.

The IEEE 123-bus reconstruction (``gigpower_b200/feeders/ieee123_allwye_noxfm_noreg.dss``, BASELINE
config C4) and the 2,541-bus radial feeder (``synth2500_radial.dss``, config C5) are not among the
reference's own test feeders, so nothing of the reference pinned them in round 1 (they were held to
the oracle only).  This script runs the reference's own ``SolutionFBS`` / ``SolutionNR3`` on them,
behind the ``opendssdirect`` stand-in, and writes ``tests/golden/ref_123.npz`` at full precision:

* 123-bus FBS: 12 seeded load-multiplier scenarios (light, heavy and random) + 4 scenarios of the
  C4 DER set (``bench.make_der_rows``) given to the reference as ``Solution.spu``;
* 123-bus NR3: 4 scenarios (nominal, light and two random multiplier vectors).  The reference holds a dense
  H of 13.6 GB (mostly untouched pages) for this feeder;
* 2,541-bus FBS: 2 scenarios (2.5 s per reference solve).

Runs only in the build container (needs ``/root/reference``).  Usage: python oracle/make_golden_123.py [--no-nr3]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                      # noqa: E402  (sets up the reference import)

sys.path.insert(0, mg.REPO)
import bench                                                  # noqa: E402  (make_der_rows: the C4 scenario set)

FEEDERS = os.path.join(mg.REPO, 'gigpower_b200', 'feeders')


def fbs_with_spu(s, spu):
    """The reference object solved again with a given ``Solution.spu`` (state reset as BASELINE.md section 3)."""
    s.V[:] = 0
    s.I[:] = 0
    s.iterations = 0
    s.Vtest = np.zeros(3, dtype=complex)
    s.spu = spu.copy()
    s.solve()
    return s


def main():
    out = {}
    z = mg.ZIPS['test_zip']
    mg.dss.TAPS = {}
    # ---------------- IEEE 123-bus, FBS ----------------
    fp = os.path.join(FEEDERS, 'ieee123_allwye_noxfm_noreg.dss')
    model = mg.read_dss(fp, {})
    rng = np.random.Generator(np.random.PCG64(123123))
    mult = rng.uniform(0.4, 1.3, size=(12, len(model.loads)))
    mult[0] = 1.0
    mult[1] = 0.25
    mult[2] = 1.5
    out['123/mult'] = mult
    V, I, it, sV, Stx, Srx = [], [], [], [], [], []
    t0 = time.time()
    for i in range(len(mult)):
        s = mg.ref_fbs(fp, z, mult[i])
        V.append(s.V.copy()), I.append(s.I.copy()), it.append(s.iterations), sV.append(s.sV.copy())
        Stx.append(s.Stx.copy()), Srx.append(s.Srx.copy())
    out['123/FBS/V'], out['123/FBS/I'], out['123/FBS/iterations'] = np.array(V), np.array(I), np.array(it)
    out['123/FBS/sV'], out['123/FBS/Stx'], out['123/FBS/Srx'] = np.array(sV), np.array(Stx), np.array(Srx)
    print('123-bus FBS multipliers', it, f'{time.time() - t0:.1f} s', flush=True)
    # C4 DER scenarios: packed rows of bench.make_der_rows -> the reference's [nb, 3] spu matrix
    s = mg.ref_fbs(fp, z)
    host = bench.HostFeeder(fp, 'fbs')
    firsts = [0, 4321, 8760 * 37 + 13, 8760 * 127 + 8759]      # scenario index = d * 8760 + h
    rows = np.concatenate([bench.make_der_rows(host.load_weights().sum(axis=1), 1, f) for f in firsts], axis=1)
    b, p = np.nonzero(host.tables.srow >= 0)
    V, I, it = [], [], []
    for j in range(rows.shape[1]):
        spu = np.zeros((len(model.bus_names), 3), dtype=complex)
        spu[b, p] = rows[host.tables.srow[b, p], j]
        fbs_with_spu(s, spu)
        V.append(s.V.copy()), I.append(s.I.copy()), it.append(s.iterations)
    out['123/der/first'], out['123/der/rows'] = np.array(firsts), rows
    out['123/der/FBS/V'], out['123/der/FBS/I'], out['123/der/FBS/iterations'] = np.array(V), np.array(I), np.array(it)
    print('123-bus FBS DER scenarios', it, flush=True)
    # ---------------- 2,541-bus, FBS ----------------
    fp5 = os.path.join(FEEDERS, 'synth2500_radial.dss')
    model5 = mg.read_dss(fp5, {})
    rng5 = np.random.Generator(np.random.PCG64(25002500))
    mult5 = rng5.uniform(0.3, 0.9, size=(2, len(model5.loads)))
    mult5[0] = 0.6
    out['2541/mult'] = mult5
    V, I, it = [], [], []
    t0 = time.time()
    for i in range(len(mult5)):
        s5 = mg.ref_fbs(fp5, z, mult5[i])
        V.append(s5.V.copy()), I.append(s5.I.copy()), it.append(s5.iterations)
        print('2,541-bus FBS', i, s5.iterations, f'{time.time() - t0:.1f} s', flush=True)
    out['2541/FBS/V'], out['2541/FBS/I'], out['2541/FBS/iterations'] = np.array(V), np.array(I), np.array(it)
    np.savez_compressed(os.path.join(mg.OUT, 'ref_123.npz'), **out)
    # ---------------- IEEE 123-bus, NR3 ----------------
    if '--no-nr3' not in sys.argv:
        X, Vn, it = [], [], []
        for i in (0, 3, 1, 5):
            t0 = time.time()
            n, its = mg.ref_nr3(fp, z, mult[i])
            X.append(n.XNR[:, 0].copy()), Vn.append(n.V.copy()), it.append(its)
            print('123-bus NR3', i, its, f'{time.time() - t0:.1f} s', flush=True)
            del n
        out['123/NR3/scenarios'] = np.array([0, 3, 1, 5])
        out['123/NR3/XNR'], out['123/NR3/V'], out['123/NR3/iterations'] = np.array(X), np.array(Vn), np.array(it)
        np.savez_compressed(os.path.join(mg.OUT, 'ref_123.npz'), **out)
    print('wrote', os.path.join(mg.OUT, 'ref_123.npz'))


if __name__ == '__main__':
    main()

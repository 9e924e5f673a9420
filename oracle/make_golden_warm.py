"""Golden vectors for warm-started solves -- TEST INFRASTRUCTURE.

SURVEY.md section 8(f)-3 (time-series driver with a warm start from the previous hour).  The
reference has no time-series driver, but its solvers do warm-start when they are simply used again:

* ``SolutionNR3.solve()`` starts from ``self.XNR`` (solution_nr3.py:598), which a second call finds at
  the previous solution; new loads go in through ``Circuit.set_kW / set_kvar`` and
  ``_init_KCL_matrices()``;
* ``SolutionFBS.solve()`` never re-initialises ``self.V`` and ``self.I`` (solution_fbs.py:76-90); with
  ``Vtest`` and ``iterations`` reset and new ``spu`` (``_set_sV_params()``) the sweep continues from them.

This script drives ONE unmodified reference object per feeder through a seeded sequence of load levels
that way and writes every step's full-precision result and iteration count to
``tests/golden/ref_warm.npz``.  Runs only in the build container.  Usage: python oracle/make_golden_warm.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                      # noqa: E402

# feeder: (seed, FBS steps, NR3 steps, taps)
CASES = {'IEEE_13_Bus_allwye_noxfm_noreg': (1313, 8, 5, {}),
         'IEEE_13_Bus_allwye': (1314, 6, 4, {'reg1': 9, 'reg2': 7, 'reg3': 9}),
         'IEEE_37_Bus_allwye_noxfm_noreg': (3737, 5, 3, {})}


def set_loads(s, base, m):
    for j, name in enumerate(s.circuit.loads.all_names()):
        s.circuit.set_kW(name, base[j][0] * m[j])
        s.circuit.set_kvar(name, base[j][1] * m[j])


def main():
    out = {}
    z = mg.ZIPS['test_zip']
    for f, (seed, n_fbs, n_nr3, taps) in CASES.items():
        fp = os.path.join(mg.FEEDERS, f + '.dss')
        mg.dss.TAPS = taps
        rng = np.random.Generator(np.random.PCG64(seed))
        s = mg.SolutionFBS(fp, zip_v=np.asarray(z))
        nload = s.circuit.loads.num_elements
        # an "hourly" drift: a common level moving by up to +-15 % per step, plus 3 % per-load noise
        level = 1.0 + np.cumsum(rng.uniform(-0.15, 0.15, size=n_fbs))
        mult = level[:, None] * (1 + rng.normal(0, 0.03, size=(n_fbs, nload)))
        mult[0] = 1.0
        out[f + '/mult'] = mult
        base = [(s.circuit.loads.get_element(n).kW, s.circuit.loads.get_element(n).kvar)
                for n in s.circuit.loads.all_names()]
        V, I, it = [], [], []
        for h in range(n_fbs):
            set_loads(s, base, mult[h])
            s._set_sV_params()
            s.Vtest = np.zeros(3, dtype=complex)
            s.iterations = 0
            s.solve()
            V.append(s.V.copy()), I.append(s.I.copy()), it.append(s.iterations)
        out[f + '/FBS/V'], out[f + '/FBS/I'], out[f + '/FBS/iterations'] = np.array(V), np.array(I), np.array(it)
        n = mg.SolutionNR3(fp, zip_v=np.asarray(z))
        X, it = [], []
        for h in range(n_nr3):
            set_loads(n, base, mult[h])
            n._init_KCL_matrices()
            mg._ft_calls[0] = 0
            n.solve()
            X.append(n.XNR[:, 0].copy()), it.append(mg._ft_calls[0])
        out[f + '/NR3/XNR'], out[f + '/NR3/iterations'] = np.array(X), np.array(it)
        print(f, 'FBS it', out[f + '/FBS/iterations'], 'NR3 it', out[f + '/NR3/iterations'], flush=True)
    np.savez_compressed(os.path.join(mg.OUT, 'ref_warm.npz'), **out)
    print('wrote', os.path.join(mg.OUT, 'ref_warm.npz'))


if __name__ == '__main__':
    main()

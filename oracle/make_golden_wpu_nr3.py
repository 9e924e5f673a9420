"""Golden vectors for a non-zero ``wpu`` in SolutionNR3 -- TEST INFRASTRUCTURE.

In the reference NR3 reads ``circuit.get_wpu_matrix()`` (all zeros, circuit.py:164-171) in two places:
``change_KCL_matrices`` writes ``b = -load * beta_S + b_factor * gamma_S - wpu`` into the constant term of the
real AND the imaginary KCL row of every bus-phase (solution_nr3.py:559-588), and ``map_output`` adds
``+ 1j * wpu`` to ``sV`` (nr3_lib/map_output.py:60).  This script gives an UNMODIFIED reference ``SolutionNR3``
object a circuit whose ``get_wpu_matrix`` returns seeded non-zero values (an attribute set on the instance; no
reference code is changed), calls the reference's own ``change_KCL_matrices(der=1e-300)`` - the smallest
argument that makes the method rewrite the constant term; 1e-300 * gamma_S vanishes in the sum, so every
other entry keeps the value ``_init_KCL_matrices`` gave it - and solves.  Output:
``tests/golden/ref_wpu_nr3.npz``.  Runs only in the build container.  Usage: python oracle/make_golden_wpu_nr3.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                      # noqa: E402

CASES = {'IEEE_13_Bus_allwye_noxfm_noreg': (1360, 6, {}), 'IEEE_13_Bus_allwye': (1361, 4, {'reg1': 9, 'reg2': 7, 'reg3': 9}),
         'IEEE_37_Bus_allwye_noxfm_noreg': (3760, 3, {})}


def main():
    out = {}
    z = mg.ZIPS['test_zip']
    for f, (seed, n, taps) in CASES.items():
        fp = os.path.join(mg.FEEDERS, f + '.dss')
        mg.dss.TAPS = taps
        rng = np.random.Generator(np.random.PCG64(seed))
        s0 = mg.SolutionNR3(fp, zip_v=np.asarray(z))
        nb = s0.circuit.buses.num_elements
        pm = np.asarray(s0.circuit.buses.get_phase_matrix('rows')).astype(bool)
        nload = s0.circuit.loads.num_elements
        mult = rng.uniform(0.6, 1.4, size=(n, nload))
        wpu = np.zeros((n, nb, 3))
        for i in range(n):
            b, p = np.nonzero(pm[1:])
            pick = rng.permutation(len(b))[:max(2, len(b) // 4)]
            wpu[i, b[pick] + 1, p[pick]] = rng.uniform(-0.05, 0.05, size=len(pick))
        out[f + '/mult'], out[f + '/wpu'] = mult, wpu
        X, V, sV, iN, it = [], [], [], [], []
        for i in range(n):
            s = mg.SolutionNR3(fp, zip_v=np.asarray(z))
            for j, name in enumerate(s.circuit.loads.all_names()):
                ld = s.circuit.loads.get_element(name)
                s.circuit.set_kW(name, ld.kW * mult[i][j])
                s.circuit.set_kvar(name, ld.kvar * mult[i][j])
            s._init_KCL_matrices()
            w = wpu[i]
            s.circuit.get_wpu_matrix = (lambda w=w, c=s.circuit: c._orient_switch(w.copy()))
            s.change_KCL_matrices(der=1e-300)
            mg._ft_calls[0] = 0
            s.solve()
            X.append(s.XNR[:, 0].copy()), V.append(s.V.copy()), sV.append(s.sV.copy()), iN.append(s.i_Node.copy())
            it.append(mg._ft_calls[0])
        out[f + '/NR3/XNR'], out[f + '/NR3/V'] = np.array(X), np.array(V)
        out[f + '/NR3/sV'], out[f + '/NR3/i_Node'], out[f + '/NR3/iterations'] = np.array(sV), np.array(iN), np.array(it)
        print(f, 'NR3 it', out[f + '/NR3/iterations'], flush=True)
    np.savez_compressed(os.path.join(mg.OUT, 'ref_wpu_nr3.npz'), **out)
    print('wrote', os.path.join(mg.OUT, 'ref_wpu_nr3.npz'))


if __name__ == '__main__':
    main()

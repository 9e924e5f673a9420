"""Golden vectors for Solution.calc_sV / calc_Vmag / calc_Stx / calc_Srx -- TEST INFRASTRUCTURE.

The four public methods (reference ``src/solution.py:177-233``) recompute the derived results from
whatever ``V``, ``I``, ``spu`` and ``wpu`` the object holds.  This script calls them on UNMODIFIED
reference ``SolutionFBS`` objects whose state has been set to seeded arbitrary values (not a
converged solution, so nothing cancels), including the per-bus form ``calc_sV(bus)``, and writes
inputs and outputs to ``tests/golden/ref_calc.npz``.

Runs only in the build container (needs ``/root/reference``).  Usage: python oracle/make_golden_calc.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                      # noqa: E402  (sets up the reference import)

CASES = ['IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_13_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg']


def main():
    out = {}
    for f in CASES:
        fp = os.path.join(mg.FEEDERS, f + '.dss')
        mg.dss.TAPS = mg.TAPS.get(f, {})
        s = mg.ref_fbs(fp, mg.ZIPS['test_zip'])
        rng = np.random.Generator(np.random.PCG64(len(f) * 1000 + 7))
        pm = np.asarray(s.phase_matrix)
        V = s.V * (1 + 0.05 * rng.standard_normal(s.V.shape)) * np.exp(1j * 0.03 * rng.standard_normal(s.V.shape))
        I = s.I + 0.01 * (rng.standard_normal(s.I.shape) + 1j * rng.standard_normal(s.I.shape)) * (s.I != 0)
        wpu = 0.02 * rng.standard_normal(s.V.shape) * pm
        spu = s.spu * rng.uniform(0.5, 1.5, size=s.spu.shape)
        s.V, s.I, s.wpu, s.spu = V.copy(), I.copy(), wpu.copy(), spu.copy()
        s.calc_sV()
        s.calc_Vmag()
        s.calc_Stx()
        s.calc_Srx()
        k = f + '/'
        out[k + 'V'], out[k + 'I'], out[k + 'wpu'], out[k + 'spu'] = V, I, wpu, spu
        out[k + 'sV'], out[k + 'Vmag'], out[k + 'Stx'], out[k + 'Srx'] = s.sV.copy(), s.Vmag.copy(), s.Stx.copy(), s.Srx.copy()
        # per-bus form: change one bus's voltage, update only its row
        name = s.circuit.buses.all_names()[3]
        bus = s.circuit.buses.get_element(name)
        i = s.circuit.buses.get_idx(bus)
        s.V[i] = s.V[i] * 0.9
        s.calc_sV(bus)
        out[k + 'bus_name'], out[k + 'bus_V'], out[k + 'bus_sV'] = np.array(name), s.V.copy(), s.sV.copy()
        print(f, 'bus', name, flush=True)
    np.savez_compressed(os.path.join(mg.OUT, 'ref_calc.npz'), **out)
    print('wrote', os.path.join(mg.OUT, 'ref_calc.npz'))


if __name__ == '__main__':
    main()

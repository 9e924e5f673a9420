"""Golden vectors for per-scenario regulator taps -- TEST INFRASTRUCTURE.

SURVEY.md section 8(f)-2: regulator tap positions as a per-scenario input.  In the reference a
tap position is fixed when the Solution object is built (``voltage_regulator.py:10-39``:
``RegControls.TapNumber()`` -> ``set_gamma``), so "scenario s with taps t_s" is by definition
"a fresh reference object built with those taps".  This script builds exactly that, one
UNMODIFIED reference ``SolutionFBS`` / ``SolutionNR3`` per scenario behind the ``opendssdirect``
stand-in, with seeded random taps in [-16, 16] per regulator and seeded load multipliers, and
writes ``tests/golden/ref_taps.npz`` at full precision.

Runs only in the build container (needs ``/root/reference``).  Usage: python oracle/make_golden_taps.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                      # noqa: E402  (sets up the reference import)

# feeder: (seed, FBS scenarios, NR3 scenarios)
CASES = {'IEEE_13_Bus_allwye': (1301, 16, 6), 'IEEE_37_Bus_allwye': (3701, 8, 3),
         'IEEE_34_Bus_allwye': (3401, 6, 0)}


def main():
    out = {}
    z = mg.ZIPS['test_zip']
    for f, (seed, n_fbs, n_nr3) in CASES.items():
        fp = os.path.join(mg.FEEDERS, f + '.dss')
        model = mg.read_dss(fp, {})
        regs = [rc.name for rc in model.regcontrols]
        rng = np.random.Generator(np.random.PCG64(seed))
        taps = rng.integers(-16, 17, size=(n_fbs, len(regs)))
        taps[0] = 16                                           # the range ends and the neutral position
        taps[1] = -16
        taps[2] = 0
        mult = rng.uniform(0.5, 1.5, size=(n_fbs, len(model.loads)))
        out[f + '/regs'] = np.array(regs)
        out[f + '/taps'] = taps
        out[f + '/mult'] = mult
        V, I, it = [], [], []
        for i in range(n_fbs):
            mg.dss.TAPS = {r: int(t) for r, t in zip(regs, taps[i])}
            s = mg.ref_fbs(fp, z, mult[i])
            V.append(s.V.copy()), I.append(s.I.copy()), it.append(s.iterations)
        out[f + '/FBS/V'], out[f + '/FBS/I'] = np.array(V), np.array(I)
        out[f + '/FBS/iterations'] = np.array(it)
        X, Vn, it = [], [], []
        for i in range(n_nr3):
            mg.dss.TAPS = {r: int(t) for r, t in zip(regs, taps[i])}
            n, its = mg.ref_nr3(fp, z, mult[i])
            X.append(n.XNR[:, 0].copy()), Vn.append(n.V.copy()), it.append(its)
        if n_nr3:
            out[f + '/NR3/XNR'], out[f + '/NR3/V'] = np.array(X), np.array(Vn)
            out[f + '/NR3/iterations'] = np.array(it)
        print(f, 'FBS it', out[f + '/FBS/iterations'], 'NR3 it', out.get(f + '/NR3/iterations'), flush=True)
    np.savez_compressed(os.path.join(mg.OUT, 'ref_taps.npz'), **out)
    print('wrote', os.path.join(mg.OUT, 'ref_taps.npz'))


if __name__ == '__main__':
    main()

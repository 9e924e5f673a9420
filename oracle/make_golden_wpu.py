"""Golden vectors for a non-zero ``Solution.wpu`` and for the volt-var droop -- TEST INFRASTRUCTURE.

``calc_sV`` (solution.py:177-220) adds ``1j * self.wpu`` - a controllable reactive injection per bus-phase -
to every bus power; ``Circuit.get_wpu_matrix`` fills it with zeros (circuit.py:164-171) and nothing in the
reference changes it, but it is a plain attribute of the Solution object.  This script sets it on an
UNMODIFIED reference ``SolutionFBS`` (seeded values on seeded bus-phases), solves, and records the result;
and it drives the reference's own ``VoltVARController.get_Q`` (volt_var_controller.py:20-38) in a
quasi-static loop around that object - solve, read |V| at the controlled bus-phases, set wpu = get_Q(|V|),
solve again from the previous state - which is the loop SURVEY.md 8(f)-3 names.  Output:
``tests/golden/ref_wpu.npz``.  Runs only in the build container.  Usage: python oracle/make_golden_wpu.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                      # noqa: E402
from gigpower.volt_var_controller import VoltVARController    # noqa: E402

CASES = {'IEEE_13_Bus_allwye_noxfm_noreg': (1350, 10, {}), 'IEEE_13_Bus_allwye': (1351, 6, {'reg1': 9, 'reg2': 7, 'reg3': 9}),
         'IEEE_37_Bus_allwye_noxfm_noreg': (3750, 5, {})}
VVC_POINTS = (0.95, 0.98, 1.02, 1.05)
VVC_ROUNDS = 4


def main():
    out = {}
    z = mg.ZIPS['test_zip']
    for f, (seed, n, taps) in CASES.items():
        fp = os.path.join(mg.FEEDERS, f + '.dss')
        mg.dss.TAPS = taps
        rng = np.random.Generator(np.random.PCG64(seed))
        s0 = mg.SolutionFBS(fp, zip_v=np.asarray(z))
        nb = s0.circuit.buses.num_elements
        pm = s0.phase_matrix.astype(bool)
        nload = s0.circuit.loads.num_elements
        mult = rng.uniform(0.6, 1.4, size=(n, nload))
        wpu = np.zeros((n, nb, 3))
        for i in range(n):
            b, p = np.nonzero(pm)
            pick = rng.permutation(len(b))[:max(2, len(b) // 4)]
            wpu[i, b[pick], p[pick]] = rng.uniform(-0.08, 0.08, size=len(pick))
        out[f + '/mult'], out[f + '/wpu'] = mult, wpu
        V, I, sV, it = [], [], [], []
        for i in range(n):
            s = mg.SolutionFBS(fp, zip_v=np.asarray(z))
            for j, name in enumerate(s.circuit.loads.all_names()):
                ld = s.circuit.loads.get_element(name)
                s.circuit.set_kW(name, ld.kW * mult[i][j])
                s.circuit.set_kvar(name, ld.kvar * mult[i][j])
            s._set_sV_params()
            s.wpu = wpu[i].copy()
            s.solve()
            V.append(s.V.copy()), I.append(s.I.copy()), sV.append(s.sV.copy()), it.append(s.iterations)
        out[f + '/FBS/V'], out[f + '/FBS/I'], out[f + '/FBS/sV'] = np.array(V), np.array(I), np.array(sV)
        out[f + '/FBS/iterations'] = np.array(it)
        # ---- volt-var droop, quasi-static: one controller per phase of two buses --------------------
        names = s0.circuit.buses.all_names()
        cand = [k for k in range(1, nb) if pm[k].all()]
        ctl_bus = [cand[len(cand) // 2], cand[-1]]
        ctl = [(k, p) for k in ctl_bus for p in range(3)]
        out[f + '/VVC/ctl'] = np.array(ctl)
        out[f + '/VVC/points'] = np.array(VVC_POINTS)
        out[f + '/VVC/q'] = np.array([-0.05, 0.05])            # minQ, maxQ in pu
        Vr, Qr, itr = [], [], []
        for i in range(min(n, 4)):
            s = mg.SolutionFBS(fp, zip_v=np.asarray(z))
            for j, name in enumerate(s.circuit.loads.all_names()):
                ld = s.circuit.loads.get_element(name)
                s.circuit.set_kW(name, ld.kW * mult[i][j])
                s.circuit.set_kvar(name, ld.kvar * mult[i][j])
            s._set_sV_params()
            vvcs = [VoltVARController(VVC_POINTS, -0.05, 0.05, names[k], p) for k, p in ctl]
            Vi, Qi, iti = [], [], []
            for rnd in range(VVC_ROUNDS):
                s.Vtest = np.zeros(3, dtype=complex)
                s.iterations = 0
                s.solve()                                   # round 0 from flat start, then from the last state
                q = np.array([v.get_Q(abs(s.V[k, p])) for v, (k, p) in zip(vvcs, ctl)])
                w = np.zeros((nb, 3))
                for (k, p), qq in zip(ctl, q):
                    w[k, p] = qq
                s.wpu = w
                Vi.append(s.V.copy()), Qi.append(q), iti.append(s.iterations)
            Vr.append(Vi), Qr.append(Qi), itr.append(iti)
        out[f + '/VVC/V'], out[f + '/VVC/Q'], out[f + '/VVC/iterations'] = np.array(Vr), np.array(Qr), np.array(itr)
        print(f, 'FBS it', out[f + '/FBS/iterations'], 'VVC it', out[f + '/VVC/iterations'].tolist(), flush=True)
    np.savez_compressed(os.path.join(mg.OUT, 'ref_wpu.npz'), **out)
    print('wrote', os.path.join(mg.OUT, 'ref_wpu.npz'))


if __name__ == '__main__':
    main()

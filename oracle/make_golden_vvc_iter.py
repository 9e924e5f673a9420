"""Golden vectors for volt-var droop INSIDE the FBS iteration -- TEST INFRASTRUCTURE.

SURVEY.md section 8(f)-3 asks for the droop of ``VoltVARController`` (volt_var_controller.py:20-38) as a
per-iteration reactive injection.  The reference ships the controller but never calls it, and its
``Solution.wpu`` - the term ``+ 1j * wpu`` of ``calc_sV`` (solution.py:203) - stays zero.  The composition
this script records is the smallest one that connects the two with the reference's own code: a SUBCLASS of
the unmodified ``SolutionFBS`` whose ``calc_sV`` first sets ``wpu[bus, phase] = controller.get_Q(abs(V[bus,
phase]))`` for every controller (the reference's own class, one object per controlled bus-phase) and then
calls the reference's ``calc_sV``.  Every evaluation of the bus powers - the one after the forward sweep,
the two per bus of the backward sweep and the final one after convergence - therefore sees the injection
that belongs to the voltage of that moment.  Output: ``tests/golden/ref_vvc_iter.npz``.
Runs only in the build container.  Usage: python oracle/make_golden_vvc_iter.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                      # noqa: E402
from gigpower.volt_var_controller import VoltVARController    # noqa: E402

CASES = {'IEEE_13_Bus_allwye_noxfm_noreg': (1370, 8, {}), 'IEEE_13_Bus_allwye': (1371, 5, {'reg1': 9, 'reg2': 7, 'reg3': 9}),
         'IEEE_37_Bus_allwye_noxfm_noreg': (3770, 4, {})}
POINTS = (0.94, 0.97, 1.0, 1.03)
MINQ, MAXQ = -0.08, 0.08


class DroopFBS(mg.SolutionFBS):
    """The reference's SolutionFBS with its own controllers consulted before every calc_sV."""
    controllers = ()

    def calc_sV(self, bus=None):
        for (k, p), ctl in self.controllers:
            self.wpu[k, p] = ctl.get_Q(abs(self.V[k, p]))
        return super().calc_sV(bus)


def main():
    out = {'points': np.array(POINTS), 'q': np.array([MINQ, MAXQ])}
    z = mg.ZIPS['test_zip']
    for f, (seed, n, taps) in CASES.items():
        fp = os.path.join(mg.FEEDERS, f + '.dss')
        mg.dss.TAPS = taps
        rng = np.random.Generator(np.random.PCG64(seed))
        s0 = mg.SolutionFBS(fp, zip_v=np.asarray(z))
        nb = s0.circuit.buses.num_elements
        pm = s0.phase_matrix.astype(bool)
        names = s0.circuit.buses.all_names()
        cand = [k for k in range(1, nb) if pm[k].all()]
        ctl = [(k, p) for k in (cand[len(cand) // 2], cand[-1]) for p in range(3)]
        single = [k for k in range(1, nb) if pm[k].sum() == 1]
        if single:
            ctl.append((single[0], int(np.nonzero(pm[single[0]])[0][0])))
        mult = rng.uniform(0.5, 1.6, size=(n, s0.circuit.loads.num_elements))
        out[f + '/ctl'], out[f + '/mult'] = np.array(ctl), mult
        V, I, sV, it, Q = [], [], [], [], []
        for i in range(n):
            s = DroopFBS(fp, zip_v=np.asarray(z))
            for j, name in enumerate(s.circuit.loads.all_names()):
                ld = s.circuit.loads.get_element(name)
                s.circuit.set_kW(name, ld.kW * mult[i][j])
                s.circuit.set_kvar(name, ld.kvar * mult[i][j])
            s._set_sV_params()
            s.controllers = [((k, p), VoltVARController(POINTS, MINQ, MAXQ, names[k], p)) for k, p in ctl]
            s.solve()
            V.append(s.V.copy()), I.append(s.I.copy()), sV.append(s.sV.copy()), it.append(s.iterations)
            Q.append(np.array([s.wpu[k, p] for k, p in ctl]))
        out[f + '/V'], out[f + '/I'], out[f + '/sV'] = np.array(V), np.array(I), np.array(sV)
        out[f + '/iterations'], out[f + '/Q'] = np.array(it), np.array(Q)
        print(f, 'it', out[f + '/iterations'], 'max|Q|', np.abs(out[f + '/Q']).max(), flush=True)
    np.savez_compressed(os.path.join(mg.OUT, 'ref_vvc_iter.npz'), **out)
    print('wrote', os.path.join(mg.OUT, 'ref_vvc_iter.npz'))


if __name__ == '__main__':
    main()

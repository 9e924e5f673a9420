"""Generate ``tests/golden/`` from the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Runs only in the build container (needs ``/root/reference``).  It imports the
reference package as it lies (``/root/reference/gigpower/src``) behind the
``opendssdirect`` stand-in, solves with the reference's own ``SolutionFBS`` /
``SolutionNR3`` and writes

* ``tests/golden/feeders/<feeder>.json``  the parsed inputs (FeederModel JSON,
  taps included) so that tests and the bench can build the same feeders on the
  GPU box, where ``/root/reference`` does not exist;
* ``tests/golden/ref_snapshots.npz``      6 feeders x 3 ZIP sets, full precision;
* ``tests/golden/ref_scenarios.npz``      seeded random load-multiplier scenarios
  on the three lines-only feeders, full precision;
* ``tests/golden/recorded_tables.json``   gigpower's own result tables recorded by
  its test-suite (``gigpower/tests/test_compare_{fbs,nr3}_dss/*.out.txt``, two
  decimals), parsed to numbers.

Usage:  python oracle/make_golden.py
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = '/root/reference/gigpower'
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(HERE, 'standin'))
sys.path.insert(0, os.path.join(REF, 'src'))

import opendssdirect as dss                                   # noqa: E402  (the stand-in)
import gigpower.solution_nr3 as ref_nr3_mod                   # noqa: E402
from gigpower.solution_fbs import SolutionFBS                 # noqa: E402
from gigpower.solution_nr3 import SolutionNR3                 # noqa: E402
from gigpower_b200.dss_reader import read_dss                 # noqa: E402

FEEDERS = os.path.join(REF, 'tests', 'test_feeders')
OUT = os.path.join(REPO, 'tests', 'golden')
ZIPS = {                                                      # test_dss_compare.py:37-40
    'constant_impedance': [1, 0, 0, 1, 0, 0, .8],
    'constant_power': [0, 0, 1, 0, 0, 1, .8],
    'test_zip': [0.10, 0.05, 0.85, 0.10, 0.05, 0.85, 0.80],
}
# regulator taps are an input: (9,7,9) reproduces all 36
# recorded tables of IEEE_13_Bus_allwye.dss; the 37-bus file sets Controlmode=OFF.
TAPS = {'IEEE_13_Bus_allwye': {'reg1': 9, 'reg2': 7, 'reg3': 9}}
ALL = ['IEEE_13_Bus_allwye', 'IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_34_Bus_allwye',
       'IEEE_34_Bus_allwye_noxfm_noreg', 'IEEE_37_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg']
SCEN = {'IEEE_13_Bus_allwye_noxfm_noreg': (13, 24, 8), 'IEEE_34_Bus_allwye_noxfm_noreg': (34, 12, 3),
        'IEEE_37_Bus_allwye_noxfm_noreg': (37, 12, 3)}       # seed, #FBS scenarios, #NR3 scenarios

_ft_calls = [0]
_orig_ft = ref_nr3_mod.compute_NR3FT


def _counting_ft(*a, **k):
    _ft_calls[0] += 1
    return _orig_ft(*a, **k)


ref_nr3_mod.compute_NR3FT = _counting_ft      # counts Newton iterations (local `itercount`, :615-627)


def ref_fbs(fp, zip_v, mult=None):
    s = SolutionFBS(fp, zip_v=np.asarray(zip_v))
    if mult is not None:
        for j, name in enumerate(s.circuit.loads.all_names()):
            ld = s.circuit.loads.get_element(name)
            s.circuit.set_kW(name, ld.kW * mult[j])
            s.circuit.set_kvar(name, ld.kvar * mult[j])
        s._set_sV_params()
    s.solve()
    return s


def ref_nr3(fp, zip_v, mult=None):
    s = SolutionNR3(fp, zip_v=np.asarray(zip_v))
    if mult is not None:
        for j, name in enumerate(s.circuit.loads.all_names()):
            ld = s.circuit.loads.get_element(name)
            s.circuit.set_kW(name, ld.kW * mult[j])
            s.circuit.set_kvar(name, ld.kvar * mult[j])
        s._init_KCL_matrices()
    _ft_calls[0] = 0
    s.solve()
    return s, _ft_calls[0]


def parse_recorded(path, algo):
    """Pull the '<param> - <algo> Results' table out of a recorded .out.txt."""
    txt = open(path).read()
    m = re.search(r'- %s Results\n(.*?)\n\s*\n' % algo, txt, flags=re.S)
    rows = {}
    for line in m.group(1).splitlines()[1:]:
        parts = line.split()
        vals = []
        for p in parts[1:]:
            c = complex(p)
            vals.append([c.real, c.imag])
        rows[parts[0]] = vals
    return rows


def main():
    os.makedirs(os.path.join(OUT, 'feeders'), exist_ok=True)
    snap, scen, recorded = {}, {}, {}
    for f in ALL:
        fp = os.path.join(FEEDERS, f + '.dss')
        taps = TAPS.get(f, {})
        dss.TAPS = taps
        model = read_dss(fp, taps)
        with open(os.path.join(OUT, 'feeders', f + '.json'), 'w') as fh:
            fh.write(model.to_json())
        for zname, z in ZIPS.items():
            s = ref_fbs(fp, z)
            k = f'{f}/{zname}/FBS/'
            for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx'):
                snap[k + p] = np.asarray(getattr(s, p))
            snap[k + 'iterations'] = np.int64(s.iterations)
            snap[k + 'convergence_diff'] = np.float64(s.convergence_diff)
            snap[k + 'converged'] = np.bool_(s.converged)
            snap[k + 'topo_order'] = np.array([s.circuit.buses.get_idx(n) for n in s.topo_order])
            n, its = ref_nr3(fp, z)
            k = f'{f}/{zname}/NR3/'
            for p in ('XNR', 'V', 'I', 'sV', 'Vmag', 'Stx', 'Srx', 'i_Node'):
                snap[k + p] = np.asarray(getattr(n, p))
            snap[k + 'iterations'] = np.int64(its)
            for algo, sub in (('FBS', 'test_compare_fbs_dss'), ('NR3', 'test_compare_nr3_dss')):
                for p in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx'):
                    path = os.path.join(REF, 'tests', sub, f'{algo}_v_DSS_{f}_{zname}-{p}.out.txt')
                    recorded[f'{f}/{zname}/{algo}/{p}'] = parse_recorded(path, algo)
            print('snapshot', f, zname, 'FBS it', s.iterations, 'NR3 it', its, flush=True)
        if f in SCEN:
            seed, n_fbs, n_nr3 = SCEN[f]
            rng = np.random.Generator(np.random.PCG64(seed))
            nload = len(model.loads)
            mult = rng.uniform(0.5, 1.5, size=(n_fbs, nload))
            mult[0] = 0.25          # light, and the two heaviest uniform cases that still converge quickly
            mult[1] = 1.5
            z = ZIPS['test_zip']
            scen[f + '/mult'] = mult
            V, I, it = [], [], []
            for i in range(n_fbs):
                s = ref_fbs(fp, z, mult[i])
                V.append(s.V.copy()), I.append(s.I.copy()), it.append(s.iterations)
            scen[f + '/FBS/V'], scen[f + '/FBS/I'] = np.array(V), np.array(I)
            scen[f + '/FBS/iterations'] = np.array(it)
            X, it = [], []
            for i in range(n_nr3):
                n, its = ref_nr3(fp, z, mult[i])
                X.append(n.XNR[:, 0].copy()), it.append(its)
            scen[f + '/NR3/XNR'], scen[f + '/NR3/iterations'] = np.array(X), np.array(it)
            print('scenarios', f, scen[f + '/FBS/iterations'], scen[f + '/NR3/iterations'], flush=True)
    np.savez_compressed(os.path.join(OUT, 'ref_snapshots.npz'), **snap)
    np.savez_compressed(os.path.join(OUT, 'ref_scenarios.npz'), **scen)
    with open(os.path.join(OUT, 'recorded_tables.json'), 'w') as fh:
        json.dump(recorded, fh, separators=(',', ':'))
    print('wrote', OUT)


if __name__ == '__main__':
    main()

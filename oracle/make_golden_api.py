"""Generate ``tests/golden/ref_api.json`` from the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Runs only in the build container (needs ``/root/reference``).  What the reference's
own ``tests/test_api.py`` exercises: the element groups of
``solution.circuit`` (names in OpenDSS order, name <-> index maps, phase matrices),
``get_nominal_bus_powers`` in both orientations for both solver classes (the reference
compares it with ``utils.get_nominal_bus_powers(dss)``, recorded here as well),
``get_params`` and the pre-solve shapes of the result matrices.

Usage:  python oracle/make_golden_api.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = '/root/reference/gigpower'
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(HERE, 'standin'))
sys.path.insert(0, os.path.join(REF, 'src'))

import opendssdirect as dss                                   # noqa: E402  (the stand-in)
from gigpower.solution_fbs import SolutionFBS                 # noqa: E402
from gigpower.solution_nr3 import SolutionNR3                 # noqa: E402
from gigpower.utils import get_nominal_bus_powers             # noqa: E402

FEEDERS = os.path.join(REF, 'tests', 'test_feeders')
OUT = os.path.join(REPO, 'tests', 'golden', 'ref_api.json')
TAPS = {'IEEE_13_Bus_allwye': {'reg1': 9, 'reg2': 7, 'reg3': 9}}       # as oracle/make_golden.py
ALL = ['IEEE_13_Bus_allwye', 'IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_34_Bus_allwye',
       'IEEE_34_Bus_allwye_noxfm_noreg', 'IEEE_37_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg']
GROUPS = ('buses', 'lines', 'loads', 'capacitors', 'transformers', 'voltage_regulators')


def cplx(a):
    a = np.asarray(a, dtype=complex)
    return {'re': a.real.tolist(), 'im': a.imag.tolist()}


def frame(df):
    return {'index': [str(x) for x in df.index], 'columns': [str(x) for x in df.columns],
            'values': cplx(df.to_numpy())}


def main():
    out = {}
    for f in ALL:
        fp = os.path.join(FEEDERS, f + '.dss')
        dss.TAPS = TAPS.get(f, {})
        rec = {}
        for cls in (SolutionFBS, SolutionNR3):
            s = cls(fp)
            r = {'orient': s._orient, 'groups': {}}
            for g in GROUPS:
                grp = getattr(s.circuit, g)
                names = list(grp.all_names())
                r['groups'][g] = {
                    'all_names': names, 'num_elements': int(grp.num_elements),
                    'idx': [int(grp.get_idx(n)) for n in names],
                    'phase_matrix_rows': np.asarray(grp.get_phase_matrix('rows')).tolist() if names else [],
                    'element_phases': {n: [str(p) for p in grp.get_element(n).phases] for n in names},
                }
            r['nominal_rows'] = frame(s.get_nominal_bus_powers(orient='rows'))
            r['nominal_cols'] = frame(s.get_nominal_bus_powers(orient='cols'))
            r['nominal_default'] = frame(s.get_nominal_bus_powers())
            p = s.get_params()
            r['params'] = {'slack_idx': int(p['slack_idx']), 'v_slack': cplx(p['v_slack']),
                           'max_iter': int(p['max_iter']), 'zip_v': [float(x) for x in p['zip_v']]}
            r['shapes'] = {q: list(np.asarray(getattr(s, q)).shape) for q in ('V', 'I', 'sV', 'Vmag', 'Stx', 'Srx')}
            rec[cls.__name__] = r
        dss.run_command('Redirect ' + fp)                     # tests/test_api.py:80-85
        dss.Solution.Solve()
        rec['dss_nominals'] = frame(get_nominal_bus_powers(dss))
        out[f] = rec
        print(f, {g: rec['SolutionFBS']['groups'][g]['num_elements'] for g in GROUPS}, flush=True)
    with open(OUT, 'w') as fh:
        json.dump(out, fh, separators=(',', ':'))
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()

"""GPU counterpart of tests/test_random_feeders.py: the CUDA kernels, through the Python mirror of the reference's
classes and the C-ABI underneath, on the sixteen seeded random radial feeders, held to the UNMODIFIED reference's own
results (tests/golden/ref_random.npz, oracle/make_golden_random.py).  Bar: 1e-9 relative on complex bus voltages,
identical iteration counts."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, ZIPS, assert_close, rel_err
from test_random_feeders import RANDOM, gold, rmodel

pytestmark = pytest.mark.gpu
TOL = 1e-9
SPECIALISED = ['rnd01', 'rnd03', 'rnd05', 'rnd09', 'rnd13']       # regulators, a transformer, the depth-first build
NR3 = [n for n in RANDOM if n + '/NR3/XNR' in gold().files]


def _path(name):
    return os.path.join(GOLDEN, 'random_feeders', name + '.json')


@pytest.mark.parametrize('name', RANDOM)
def test_fbs_on_random_feeders_matches_the_reference(name):
    """solve() (the snapshot of the feeder as written) and solve_batch (three load-multiplier scenarios) on the
    streaming kernel in both step orders; on five of the feeders also on the feeder-specialised kernel."""
    from gigpower_b200 import _cabi
    from gigpower_b200.solution_fbs import SolutionFBS
    g = gold()
    s = SolutionFBS(_path(name), zip_v=np.asarray(ZIPS['test_zip']))
    s.solve()
    assert s.iterations == int(g[name + '/FBS/iterations'][0]) and s.converged
    for p in ('V', 'I', 'sV', 'Stx', 'Srx'):
        assert_close(getattr(s, p), g[f'{name}/FBS/{p}'][0], TOL, p)
    mult = g[name + '/mult']
    try:
        for order in (1, 0):
            _cabi.check(_cabi.gp_set_option(b'fbs_order', order))
            r = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'sV', 'Stx', 'Srx'))
            assert list(r.iterations) == list(g[name + '/FBS/iterations']) and (r.status == 0).all()
            for p in ('V', 'I', 'sV', 'Stx', 'Srx'):
                assert_close(getattr(r, p), g[f'{name}/FBS/{p}'], TOL, p)
    finally:
        _cabi.gp_set_option(b'fbs_order', 1)
    if name in SPECIALISED:
        assert s.specialise() and _cabi.gp_fbs_active_kernel(s._feeder.handle) == 5
        r = s.solve_batch(load_mult=mult, outputs=('V', 'I', 'Vmag', 'loss'))
        assert list(r.iterations) == list(g[name + '/FBS/iterations']) and (r.status == 0).all()
        assert_close(r.V, g[name + '/FBS/V'], TOL, 'V')
        assert_close(r.I, g[name + '/FBS/I'], TOL, 'I')
        assert rel_err(r.Vmag, np.abs(g[name + '/FBS/V'])) <= TOL
        loss = (g[name + '/FBS/Stx'] - g[name + '/FBS/Srx']).sum(axis=(1, 2))
        assert rel_err(r.losses, loss) <= TOL


@pytest.mark.parametrize('name', NR3)
def test_nr3_on_random_feeders_matches_the_reference(name):
    """Both NR3 kernels (the phase-lane kernel on the feeders without regulators) against the reference's X and
    Newton iteration counts."""
    from gigpower_b200 import _cabi
    from gigpower_b200.solution_nr3 import SolutionNR3
    g = gold()
    s = SolutionNR3(_path(name), zip_v=np.asarray(ZIPS['test_zip']))
    X = g[name + '/NR3/XNR']
    mult = g[name + '/mult'][:len(X)]
    kernels = (1,) if rmodel(name).regcontrols else (1, 2)
    try:
        for kern in kernels:
            _cabi.check(_cabi.gp_set_option(b'nr3_kernel', kern))
            r = s.solve_batch(load_mult=mult, outputs=('X', 'V'))
            assert list(r.iterations) == list(g[name + '/NR3/iterations']), kern
            assert (r.status == 0).all()
            assert rel_err(r.XNR, X) <= TOL, kern
            assert rel_err(r.V, g[name + '/NR3/V']) <= TOL, kern
    finally:
        _cabi.gp_set_option(b'nr3_kernel', 0)

"""Shared helpers for the test-suite (feeder fixtures, ZIP sets, scenarios)."""
import json
import os

import numpy as np

from gigpower_b200.dss_reader import read_dss

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, 'tests', 'golden')

ZIPS = {   # reference tests/test_dss_compare.py:37-40
    'constant_impedance': [1, 0, 0, 1, 0, 0, .8],
    'constant_power': [0, 0, 1, 0, 0, 1, .8],
    'test_zip': [0.10, 0.05, 0.85, 0.10, 0.05, 0.85, 0.80],
}
ALL_FEEDERS = ['IEEE_13_Bus_allwye', 'IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_34_Bus_allwye',
               'IEEE_34_Bus_allwye_noxfm_noreg', 'IEEE_37_Bus_allwye', 'IEEE_37_Bus_allwye_noxfm_noreg']
LINES_ONLY = [f for f in ALL_FEEDERS if f.endswith('noxfm_noreg')]


def feeder_path(name):
    return os.path.join(GOLDEN, 'feeders', name + '.json')


def load_model(name):
    return read_dss(feeder_path(name))


_cache = {}


def snapshots():
    if 'snap' not in _cache:
        _cache['snap'] = np.load(os.path.join(GOLDEN, 'ref_snapshots.npz'))
    return _cache['snap']


def scenarios():
    if 'scen' not in _cache:
        _cache['scen'] = np.load(os.path.join(GOLDEN, 'ref_scenarios.npz'))
    return _cache['scen']


def recorded():
    if 'rec' not in _cache:
        with open(os.path.join(GOLDEN, 'recorded_tables.json')) as f:
            _cache['rec'] = json.load(f)
    return _cache['rec']


def rel_err(a, b):
    """max |a-b| / max(1, |b|) elementwise -- the 1e-9 'relative on complex bus
    voltages' bar of BASELINE.json (voltages are O(1) pu, currents O(1))."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def true_rel_err(a, b, floor=1e-6):
    """max |a-b| / |b| over the entries with |b| > floor: a STRICTLY relative error (rel_err above is
    absolute below 1 pu, which is 100x looser than "relative" for currents and powers of ~1e-2 pu).
    Entries at or below the floor (structural zeros, absent phases) are held to an absolute floor * tol
    by the callers through ``small_abs_err``."""
    a, b = np.asarray(a), np.asarray(b)
    m = np.abs(b) > floor
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m]))) if m.any() else 0.0


def small_abs_err(a, b, floor=1e-6):
    """max |a-b| over the entries with |b| <= floor."""
    a, b = np.asarray(a), np.asarray(b)
    m = np.abs(b) <= floor
    return float(np.max(np.abs(a[m] - b[m]))) if m.any() else 0.0


def assert_close(a, b, tol=1e-9, what='', small=1e-13):
    """The parity bar on one result array, three ways: rel_err (absolute below 1 pu, the bar of BASELINE.json on
    voltages), STRICTLY relative above 1e-6 pu (currents and powers of 1e-2 pu are not let off 100x easier) and
    absolute on the entries below that floor (structural zeros, absent phases)."""
    assert rel_err(a, b) <= tol, (what, 'rel_err', rel_err(a, b))
    assert true_rel_err(a, b) <= tol, (what, 'true_rel_err', true_rel_err(a, b))
    assert small_abs_err(a, b) <= small, (what, 'small_abs_err', small_abs_err(a, b))


def ref_123():
    """tests/golden/ref_123.npz (oracle/make_golden_123.py): the unmodified reference's FBS and NR3 on
    the authored IEEE 123-bus feeder and its FBS on the 2,541-bus feeder."""
    if 'r123' not in _cache:
        _cache['r123'] = np.load(os.path.join(GOLDEN, 'ref_123.npz'))
    return _cache['r123']


def taps_golden():
    """tests/golden/ref_taps.npz (oracle/make_golden_taps.py): one unmodified reference object
    per scenario, built with that scenario's regulator taps."""
    if 'taps' not in _cache:
        _cache['taps'] = np.load(os.path.join(GOLDEN, 'ref_taps.npz'))
    return _cache['taps']


def pinned_34():
    """tests/golden/ref_34_taps.npz (oracle/pin_taps_34.py): the taps behind gigpower's recorded
    tables of IEEE_34_Bus_allwye.dss, and the unmodified reference's full-precision results at them."""
    if 'p34' not in _cache:
        _cache['p34'] = np.load(os.path.join(GOLDEN, 'ref_34_taps.npz'))
    return _cache['p34']


def pinned_34_taps():
    z = pinned_34()
    return {str(n): int(t) for n, t in zip(z['regcontrols'], z['taps'])}


TAP_FEEDERS = ['IEEE_13_Bus_allwye', 'IEEE_37_Bus_allwye', 'IEEE_34_Bus_allwye']


def warm_golden():
    """tests/golden/ref_warm.npz (oracle/make_golden_warm.py): one unmodified reference object per
    feeder driven through a sequence of load levels, each solve starting from the previous one."""
    if 'warm' not in _cache:
        _cache['warm'] = np.load(os.path.join(GOLDEN, 'ref_warm.npz'))
    return _cache['warm']


WARM_FEEDERS = {'IEEE_13_Bus_allwye_noxfm_noreg': {}, 'IEEE_13_Bus_allwye': {'reg1': 9, 'reg2': 7, 'reg3': 9},
                'IEEE_37_Bus_allwye_noxfm_noreg': {}}


def wpu_golden():
    """tests/golden/ref_wpu.npz (oracle/make_golden_wpu.py): unmodified reference objects with a non-zero
    Solution.wpu attribute, and the reference's VoltVARController driven around them."""
    if 'wpu' not in _cache:
        _cache['wpu'] = np.load(os.path.join(GOLDEN, 'ref_wpu.npz'))
    return _cache['wpu']

"""The host-side mirror of gigpower's Python API against facts recorded from the unmodified
reference (tests/golden/ref_api.json, oracle/make_golden_api.py) -- what the reference's own
tests/test_api.py exercises, minus the solve (constructors need no GPU)."""
import json
import os

import numpy as np
import pytest

from helpers import ALL_FEEDERS, GOLDEN, feeder_path

GROUPS = ('buses', 'lines', 'loads', 'capacitors', 'transformers', 'voltage_regulators')


@pytest.fixture(scope='module')
def api():
    with open(os.path.join(GOLDEN, 'ref_api.json')) as fh:
        return json.load(fh)


def _cls(name):
    from gigpower_b200.solution_fbs import SolutionFBS
    from gigpower_b200.solution_nr3 import SolutionNR3
    return {'SolutionFBS': SolutionFBS, 'SolutionNR3': SolutionNR3}[name]


def _values(rec):
    return np.array(rec['values']['re']) + 1j * np.array(rec['values']['im'])


def _same_frame(df, rec):
    assert [str(x) for x in df.index] == rec['index']
    assert [str(x) for x in df.columns] == rec['columns']
    want = _values(rec)
    assert df.to_numpy().shape == want.shape
    assert np.abs(df.to_numpy() - want).max() <= 1e-9 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize('cls', ['SolutionFBS', 'SolutionNR3'])
@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_element_groups_match_the_reference(api, feeder, cls):
    rec = api[feeder][cls]
    s = _cls(cls)(feeder_path(feeder))
    assert s._orient == rec['orient']
    for g in GROUPS:
        grp, want = getattr(s.circuit, g), rec['groups'][g]
        names = list(grp.all_names())
        assert names == want['all_names'], g
        assert grp.num_elements == want['num_elements'], g
        assert [grp.get_idx(n) for n in names] == want['idx'], g
        assert [grp.get_name(i) for i in want['idx']] == names, g
        if names:
            assert np.array_equal(np.asarray(grp.get_phase_matrix('rows')), np.array(want['phase_matrix_rows'])), g
            assert np.array_equal(np.asarray(grp.get_phase_matrix('cols')), np.array(want['phase_matrix_rows']).T), g
        for n in names:
            e = grp.get_element(n)
            assert e.__name__ == n and grp.get_element(grp.get_idx(n)) is e
            assert [str(p) for p in e.phases] == want['element_phases'][n], (g, n)
        assert len(list(grp.get_elements())) >= len(names)
        with pytest.raises(KeyError):                         # circuit_element_group.py:76-87
            grp.get_element('no such element')
        with pytest.raises(KeyError):
            grp.get_element(3.5)


@pytest.mark.parametrize('cls', ['SolutionFBS', 'SolutionNR3'])
@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_nominal_bus_powers_match_the_reference(api, feeder, cls):
    """tests/test_api.py:73-87: Circuit's nominal powers equal those read from OpenDSS."""
    rec = api[feeder][cls]
    s = _cls(cls)(feeder_path(feeder))
    _same_frame(s.get_nominal_bus_powers(orient='rows'), rec['nominal_rows'])
    _same_frame(s.get_nominal_bus_powers(orient='cols'), rec['nominal_cols'])
    _same_frame(s.get_nominal_bus_powers(), rec['nominal_default'])
    if cls == 'SolutionFBS':      # the reference's own assertion (its NR3 frame comes out transposed)
        _same_frame(s.get_nominal_bus_powers(orient='rows'), api[feeder]['dss_nominals'])


@pytest.mark.parametrize('cls', ['SolutionFBS', 'SolutionNR3'])
@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_params_and_result_shapes_match_the_reference(api, feeder, cls):
    rec = api[feeder][cls]
    s = _cls(cls)(feeder_path(feeder))
    p = s.get_params()
    assert int(p['slack_idx']) == rec['params']['slack_idx']
    assert int(p['max_iter']) == rec['params']['max_iter']
    assert np.allclose(np.asarray(p['v_slack']), _values({'values': rec['params']['v_slack']}), rtol=0, atol=1e-15)
    assert [float(x) for x in p['zip_v']] == rec['params']['zip_v']
    for q, shape in rec['shapes'].items():
        assert list(np.asarray(getattr(s, q)).shape) == shape, q


def test_out_of_scope_entry_points_fail_with_an_explanation():
    """N10 change_KCL_matrices is outside the accelerated path (its inputs are per-scenario arrays of solve_batch): a
    caller gets a clear error.  SolutionDSS - OpenDSS itself behind gigpower's accessors - is not part of this package
    at all: there is no such module."""
    s = _cls('SolutionNR3')(feeder_path(ALL_FEEDERS[1]))
    with pytest.raises(NotImplementedError, match='solve_batch'):
        s.change_KCL_matrices(der=0.1)
    with pytest.raises(ModuleNotFoundError):
        import gigpower_b200.solution_dss  # noqa: F401


def test_solution_keeps_the_reference_method_names():
    """The public methods of the reference's Solution classes a caller may invoke (solution.py:127-315,
    solution_fbs.py, solution_nr3.py) exist on the drop-in classes."""
    fbs, nr3 = _cls('SolutionFBS'), _cls('SolutionNR3')
    for name in ('solve', 'get_V', 'get_Vmag', 'get_I', 'get_Stx', 'get_Srx', 'get_sV', 'get_data_frame',
                 'get_nominal_bus_powers', 'get_params', 'print_solution', 'set_zip_values'):
        assert callable(getattr(fbs, name)) and callable(getattr(nr3, name)), name
    for name in ('calc_sV', 'calc_Vmag', 'calc_Stx', 'calc_Srx'):           # solution.py:177-233
        assert callable(getattr(fbs, name)), name
    for name in ('map_XNR', 'change_KCL_matrices'):
        assert callable(getattr(nr3, name)), name

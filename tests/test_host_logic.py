"""CPU tests of the host side: reader, circuit mirror, flattener, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from gigpower_b200.circuit import Circuit
from gigpower_b200.flatten import flatten_fbs, topo_sort
from oracle.gp_oracle import OracleCircuit
from helpers import ALL_FEEDERS, LINES_ONLY, ZIPS, REPO, load_model, snapshots


@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_circuit_matches_oracle_extraction(feeder):
    model = load_model(feeder)
    Circuit._set_zip_values(ZIPS['test_zip'])
    c, o = Circuit(model), OracleCircuit(model, ZIPS['test_zip'])
    assert c.buses.all_names() == o.bus_names
    assert np.array_equal(c.buses.get_phase_matrix('rows'), o.bus_pm)
    assert np.array_equal(np.array([e.FZpu for e in c.lines.get_elements()]), o.FZpu)
    assert np.array_equal(np.array([e.rmat for e in c.lines.get_elements()]), o.rmat)
    assert np.array_equal(c.get_spu_matrix(), o.spu_matrix())
    assert np.array_equal(c.get_cappu_matrix(), o.cappu)
    assert np.array_equal(c.get_aPQ_matrix(), o.aPQ)
    assert np.array_equal(c.get_tx_idx_matrix()[:o.nl], o.line_tx)
    assert c.get_total_lines() == o.nl + o.ntf + o.nvr


@pytest.mark.parametrize('feeder', ALL_FEEDERS)
def test_topo_order_matches_reference(feeder):
    model = load_model(feeder)
    t = flatten_fbs(Circuit(model))
    want = snapshots()[f'{feeder}/test_zip/FBS/topo_order']
    assert [model.bus_names.index(n) for n in t.topo_order] == list(want)
    assert abs(t.tolerance - 1e-6) < 1e-20


def test_packed_row_counts_match_survey_table():
    # P_b, P_l, P_L of the three lines-only feeders
    want = {'IEEE_13_Bus_allwye_noxfm_noreg': (38, 35, 19), 'IEEE_34_Bus_allwye_noxfm_noreg': (89, 86, 54),
            'IEEE_37_Bus_allwye_noxfm_noreg': (117, 114, 55)}
    for f, (pb, pl, pL) in want.items():
        t = flatten_fbs(Circuit(load_model(f)))
        assert (t.n_vrows, t.n_irows, t.n_srows) == (pb, pl, pL)
        assert t.n_steps == len(t.topo_order) - 1
        assert t.child_count.sum() == t.child_irow.shape[0]


def test_topo_sort_cycle_raises():
    with pytest.raises(ValueError, match='Not radial'):
        topo_sort(['a', 'b', 'c'], {'a': ['b'], 'b': ['c'], 'c': ['a']})


def test_set_kw_changes_spu_like_reference():
    c = Circuit(load_model(LINES_ONLY[0]))
    before = c.get_spu_matrix().copy()
    c.set_kW('671', 2 * 1155)
    after = c.get_spu_matrix()
    b = c.buses.get_idx('671')
    assert np.allclose(after[b].real, 2 * before[b].real) and np.allclose(after[b].imag, before[b].imag)
    with pytest.raises(KeyError):
        c.loads.get_element(3.5)


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(REPO, 'include', 'gigpower_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = sorted(set(re.findall(r'\b(gp_[a-z0-9_]+)\s*\(', hdr)))
    assert declared, 'no declarations found'
    lib = ctypes.CDLL(os.path.join(REPO, 'gigpower_b200', 'libgigpower_b200.so'))
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    lib.gp_version.restype = ctypes.c_int
    assert lib.gp_version() >= 100


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, 'gigpower_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f


@pytest.mark.parametrize('feeder,warps,ig', [('IEEE_13_Bus_allwye_noxfm_noreg', 6, 0), ('IEEE_13_Bus_allwye', 8, 1),
                                             ('IEEE_37_Bus_allwye', 1, 0), ('IEEE_34_Bus_allwye_noxfm_noreg', 4, 1)])
def test_specialised_kernel_source_generates_and_compiles(feeder, warps, ig):
    """Host-only half of gp_fbs_specialise: the generator writes the feeder's sweep schedule as
    straight-line CUDA and NVRTC compiles it for sm_100a (no GPU needed for either)."""
    import ctypes as C
    from gigpower_b200 import _cabi
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.flatten import flatten_fbs
    t = flatten_fbs(Circuit(load_model(feeder)))
    d, keep = _cabi.make_fbs_desc(t)
    n = _cabi.gp_fbs_jit_source(C.byref(d), warps, ig, None, 0)
    assert n > 0
    buf = C.create_string_buffer(n + 1)
    assert _cabi.gp_fbs_jit_source(C.byref(d), warps, ig, buf, n + 1) == n
    src = buf.value.decode()
    assert len(src) == n and 'gp_fbs_jit_kernel' in src and f'#define NV {t.n_vrows}' in src
    # one forward and one backward block per step, in sweep order
    assert src.count('// forward step') == t.n_steps and src.count('// backward step') == t.n_steps
    assert src.index('// forward step 0\n') < src.index(f'// backward step {t.n_steps - 1}\n') \
        < src.index('// backward step 0\n')
    size = C.c_int64(0)
    _cabi.check(_cabi.gp_jit_compile(buf.value, None, 0, C.byref(size)))
    assert size.value > 10000
    cubin = C.create_string_buffer(size.value)
    _cabi.check(_cabi.gp_jit_compile(buf.value, cubin, size.value, C.byref(size)))
    assert cubin.raw[:4] == b'\x7fELF'
    # bad arguments are errors, not crashes
    assert _cabi.gp_fbs_jit_source(C.byref(d), 0, ig, None, 0) < 0
    assert _cabi.gp_jit_compile(b'this is not CUDA', None, 0, C.byref(size)) == -4


@pytest.mark.parametrize('feeder,n_rows', [('IEEE_13_Bus_allwye', 3), ('IEEE_34_Bus_allwye', 6), ('IEEE_37_Bus_allwye', 3),
                                           ('IEEE_13_Bus_allwye_noxfm_noreg', 0)])
def test_gamma_rows_follow_the_step_table(feeder, n_rows):
    """Rows of a per-scenario regulator-ratio array (GpSolveExtras.gamma_dev): one per regulated
    (step, phase) in step then phase order, each owned by the regulator that set that ratio."""
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.flatten import flatten_fbs, flatten_nr3, EDGE_VR
    for taps in ({}, {'reg1': 5, 'reg2': -7, 'reg3': 11, 'reg1a': 1, 'reg1b': 2, 'reg1c': 3, 'reg2a': 4,
                      'reg2b': 5, 'reg2c': 6}):
        model = load_model(feeder)
        for rc in model.regcontrols:
            rc.tap = taps.get(rc.name, rc.tap)
        c = Circuit(model)
        vrs = c.voltage_regulators.get_elements()
        for t in (flatten_fbs(c), flatten_nr3(c)):
            assert len(t.grow_vr) == n_rows
            rows = t.grow[t.grow >= 0]
            assert list(rows) == list(range(n_rows))                 # step-major, phase-minor numbering
            assert ((t.grow >= 0) == ((t.kind[:, None] == EDGE_VR) & (t.vr_gamma != 0))).all()
            for i, p in zip(*np.nonzero(t.grow >= 0)):
                assert t.vr_gamma[i, p] == vrs[t.grow_vr[t.grow[i, p]]].gamma


def test_packaged_feeders_are_the_golden_ones():
    """bench.py and smoke() read their input feeders from gigpower_b200/feeders; the three IEEE ones are
    copies of the parsed reference feeders under tests/golden/feeders."""
    import filecmp
    import os
    from helpers import REPO, GOLDEN
    for f in ('IEEE_13_Bus_allwye_noxfm_noreg', 'IEEE_34_Bus_allwye_noxfm_noreg', 'IEEE_37_Bus_allwye_noxfm_noreg'):
        assert filecmp.cmp(os.path.join(REPO, 'gigpower_b200', 'feeders', f + '.json'),
                           os.path.join(GOLDEN, 'feeders', f + '.json'), shallow=False)

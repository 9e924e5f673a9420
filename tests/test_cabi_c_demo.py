"""The C-ABI from plain C (examples/cabi_fbs_demo.c): generate the feeder tables of the IEEE 13-bus
feeder as a C header, compile the demo with gcc against include/gigpower_b200.h and the in-tree
library, run it.  Without a CUDA device the program must fail loudly (exit 3, no CPU fallback);
on a GPU its printed voltages are held to the oracle at 1e-9 with equal iteration counts."""
import os
import subprocess

import numpy as np
import pytest

from helpers import REPO, ZIPS, feeder_path, rel_err

FEEDER, S = 'IEEE_13_Bus_allwye', 5


def _c_array(ctype, name, a):
    flat = np.asarray(a).reshape(-1)
    body = ', '.join(repr(float(x)) if ctype == 'double' else str(int(x)) for x in flat) or '0'
    return f'static const {ctype} {name}[{max(flat.size, 1)}] = {{{body}}};\n'


def _build(tmp_path):
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.dss_reader import read_dss
    from gigpower_b200.flatten import flatten_fbs, load_weights
    from gigpower_b200.solution import Solution
    Solution.set_zip_values(np.asarray(ZIPS['test_zip']))
    model = read_dss(feeder_path(FEEDER), {'reg1': 9, 'reg2': 7, 'reg3': 9})
    circuit = Circuit(model)
    t = flatten_fbs(circuit)

    mult = np.random.Generator(np.random.PCG64(99)).uniform(0.5, 1.5, size=(S, circuit.loads.num_elements))
    rows = np.ascontiguousarray(load_weights(circuit, t.srow, t.n_srows) @ mult.T)                 # [n_srows, S] complex
    hdr = [f'#define GP_DEMO_S {S}\n', f'#define GP_DEMO_N_STEPS {t.n_steps}\n', f'#define GP_DEMO_N_VROWS {t.n_vrows}\n',
           f'#define GP_DEMO_N_IROWS {t.n_irows}\n', f'#define GP_DEMO_N_SROWS {t.n_srows}\n',
           f'#define GP_DEMO_N_CHILD {t.child_irow.shape[0]}\n', f'#define GP_DEMO_N_LINES {t.n_lines}\n',
           f'#define GP_DEMO_TOL {float(t.tolerance)!r}\n']
    for name in ('bus_vrow', 'par_vrow', 'edge_irow', 'load_srow', 'kind', 'child_begin', 'child_count', 'child_irow',
                 'root_vrow', 'root_load_srow', 'line_irow', 'line_tx_vrow', 'line_rx_vrow'):
        hdr.append(_c_array('int32_t', 'gp_demo_' + name, getattr(t, name)))
    for name in ('z', 'cappu', 'vr_gamma', 'root_cappu', 'vref', 'zip'):
        hdr.append(_c_array('double', 'gp_demo_' + name, getattr(t, name)))
    hdr.append(_c_array('double', 'gp_demo_spu', np.stack([rows.real, rows.imag], axis=-1)))
    (tmp_path / 'feeder_data.h').write_text('#include <stdint.h>\n' + ''.join(hdr))
    exe = tmp_path / 'cabi_fbs_demo'
    lib = os.path.join(REPO, 'gigpower_b200')
    cmd = ['gcc', '-O2', '-Wall', '-Werror', '-I', os.path.join(REPO, 'include'), '-I', str(tmp_path),
           os.path.join(REPO, 'examples', 'cabi_fbs_demo.c'), '-L', lib, '-lgigpower_b200', f'-Wl,-rpath,{lib}',
           '-o', str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe, model, mult, t


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason='needs a box without a GPU')
def test_c_program_links_and_fails_loudly_without_cuda(tmp_path):
    exe, *_ = _build(tmp_path)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert 'libgigpower_b200 version' in res.stdout
    assert res.returncode == 3 and 'gp_fbs_create failed' in res.stderr     # no CPU fallback


@pytest.mark.gpu
def test_c_program_matches_the_oracle(tmp_path):
    from oracle.gp_oracle import FBSOracle
    exe, model, mult, t = _build(tmp_path)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    o = FBSOracle(model, ZIPS['test_zip'])
    want = o.solve(o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult))
    its, V = np.zeros(S, int), np.zeros((t.n_vrows, S), complex)
    for line in res.stdout.splitlines():
        w = line.split()
        if w[0] == 'scenario':
            its[int(w[1])] = int(w[3])
            assert int(w[5]) == 0
        elif w[0] == 'V':
            V[int(w[1]), int(w[2])] = float(w[3]) + 1j * float(w[4])
    assert np.array_equal(its, want['iterations'])
    b, p = np.nonzero(t.vrow >= 0)
    assert rel_err(V[t.vrow[b, p]].T, want['V'][:, b, p]) <= 1e-9

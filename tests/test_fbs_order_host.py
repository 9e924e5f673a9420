"""The streaming FBS kernel's depth-first step order and its dropped (dead) parent-voltage stores, checked
WITHOUT a GPU: ``gp_fbs_schedule`` (host only) returns the event list and the per-step flag words the
kernel is given, and this test runs the kernel's loop over exactly those tables in NumPy - batched over
scenarios, one forward or backward step per event, a single V array - and holds the result to the oracle
(the reference's two full sweeps): equal iteration counts, V and I to 1e-12.  Test infrastructure only."""
import ctypes as C

import numpy as np
import pytest

from helpers import ALL_FEEDERS, ZIPS, load_model, REPO


def _schedule(t):
    from gigpower_b200 import _cabi
    d, keep = _cabi.make_fbs_desc(t)
    ev = np.zeros(4 * max(t.n_steps, 1), dtype=np.int32)
    kind = np.zeros(max(t.n_steps, 1), dtype=np.int32)
    n = _cabi.gp_fbs_schedule(C.byref(d), ev.ctypes.data_as(_cabi.i32p), kind.ctypes.data_as(_cabi.i32p))
    assert n == 2 * t.n_steps, 'the feeder must give a depth-first schedule'
    return ev[:2 * n].reshape(n, 2), kind[:t.n_steps]


def emulate(t, rows, events, kind, order, maxiter=100):
    """The kernel's solve loop (fbs_solve_kernel) in NumPy.  rows: packed spu [n_srows, S]."""
    S = rows.shape[1]
    V = np.zeros((t.n_vrows, S), dtype=complex)
    I = np.zeros((max(t.n_irows, 1), S), dtype=complex)
    z = np.asarray(t.z).reshape(t.n_steps, 3, 3, 2)
    Z = z[..., 0] + 1j * z[..., 1]
    vref = np.asarray(t.vref).reshape(3, 2)
    vref = vref[:, 0] + 1j * vref[:, 1]
    aZ, aI, aPQ = 0, 0, 0
    aPQ, aI, aZ = [float(x) for x in t.zip]
    active = np.ones(S, dtype=bool)
    its = np.zeros(S, dtype=int)
    gam = np.asarray(t.vr_gamma).reshape(t.n_steps, 3)
    cap = np.asarray(t.cappu).reshape(t.n_steps, 3)

    def fwd(i, a):
        if (kind[i] & 0xff) == 1:                       # GP_EDGE_VR
            for p in range(3):
                if gam[i, p] != 0 and t.bus_vrow[i, p] >= 0:
                    vp = V[t.par_vrow[i, p], a] if t.par_vrow[i, p] >= 0 else 0
                    V[t.bus_vrow[i, p], a] = gam[i, p] * vp
            return
        for p in range(3):
            if t.bus_vrow[i, p] < 0:
                continue
            v = V[t.par_vrow[i, p], a].copy() if t.par_vrow[i, p] >= 0 else np.zeros(a.sum(), dtype=complex)
            for q in range(3):
                if t.edge_irow[i, q] >= 0:
                    v = v - Z[i, p, q] * I[t.edge_irow[i, q], a]
            V[t.bus_vrow[i, p], a] = v

    def bwd(i, a):
        dead = (kind[i] >> 19) & 7
        if (kind[i] & 0xff) == 1:
            for p in range(3):
                if gam[i, p] != 0 and t.par_vrow[i, p] >= 0 and not (dead >> p) & 1:
                    vb = V[t.bus_vrow[i, p], a] if t.bus_vrow[i, p] >= 0 else 0
                    V[t.par_vrow[i, p], a] = (1.0 / gam[i, p]) * vb
            return
        n_act = int(a.sum())
        In = np.zeros((3, n_act), dtype=complex)
        Vb = np.zeros((3, n_act), dtype=complex)
        for p in range(3):
            if t.bus_vrow[i, p] < 0:
                continue
            v = V[t.bus_vrow[i, p], a]
            Vb[p] = v
            mag = np.abs(v)
            sp = rows[t.load_srow[i, p], a] if t.load_srow[i, p] >= 0 else 0
            sv = sp * (aPQ + aI * mag + aZ * mag ** 2) - 1j * cap[i, p] * mag ** 2
            with np.errstate(all='ignore'):
                In[p] = np.where(v != 0, np.conj(sv / np.where(v != 0, v, 1)), 0)
        for c in range(t.child_begin[i], t.child_begin[i] + t.child_count[i]):
            for p in range(3):
                if t.child_irow[c, p] >= 0:
                    In[p] = In[p] + I[t.child_irow[c, p], a]
        for p in range(3):
            if t.edge_irow[i, p] >= 0:
                In[p] = np.nan_to_num(In[p])
                I[t.edge_irow[i, p], a] = In[p]
            else:
                In[p] = 0
        for p in range(3):
            if t.bus_vrow[i, p] < 0 or t.par_vrow[i, p] < 0 or (dead >> p) & 1:
                continue
            v = Vb[p].copy()
            for q in range(3):
                if t.edge_irow[i, q] >= 0:
                    v = v + Z[i, p, q] * In[q]
            V[t.par_vrow[i, p], a] = v

    for _ in range(maxiter):
        if not active.any():
            break
        a = active.copy()
        for p in range(3):
            V[t.root_vrow[p], a] = vref[p]
        if order == 'dfs':
            for step, fl in events:
                (bwd if fl & 1 else fwd)(int(step), a)
        else:
            for i in range(t.n_steps):
                fwd(i, a)
            for i in range(t.n_steps - 1, -1, -1):
                bwd(i, a)
        its[a] += 1
        diff = np.max(np.abs(V[np.asarray(t.root_vrow)][:, a] - vref[:, None]), axis=0)
        done = np.zeros(S, dtype=bool)
        done[a] = diff <= t.tolerance
        active &= ~done
    return V, I, its


import os                                                    # noqa: E402
AUTHORED = os.path.join(REPO, 'gigpower_b200', 'feeders', 'ieee123_allwye_noxfm_noreg.dss')


@pytest.mark.parametrize('feeder', ALL_FEEDERS + ['ieee123'])
@pytest.mark.parametrize('order', ['dfs', 'sweeps'])
def test_kernel_loop_over_the_library_schedule_matches_the_oracle(feeder, order):
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.dss_reader import read_dss
    from gigpower_b200.flatten import flatten_fbs
    from gigpower_b200.solution import Solution
    from oracle.gp_oracle import FBSOracle
    Solution.set_zip_values(np.asarray(ZIPS['test_zip']))
    model = read_dss(AUTHORED) if feeder == 'ieee123' else load_model(feeder)
    t = flatten_fbs(Circuit(model))
    events, kind = _schedule(t)
    # every step goes down once and comes up once; a step's events bracket its children's
    assert sorted(events[events[:, 1] & 1 == 0, 0]) == list(range(t.n_steps))
    assert sorted(events[events[:, 1] & 1 == 1, 0]) == list(range(t.n_steps))
    o = FBSOracle(model, ZIPS['test_zip'])
    rng = np.random.Generator(np.random.PCG64(5))
    S = 24
    mult = rng.uniform(0.4, 1.5, size=(S, len(o.c.load_names)))
    mult[0] = 1.0
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    want = o.solve(spu)
    b, p = np.nonzero(t.srow >= 0)
    rows = np.zeros((max(t.n_srows, 1), S), dtype=complex)
    rows[t.srow[b, p]] = spu[:, b, p].T
    V, I, its = emulate(t, rows, events, kind, order)
    assert list(its) == list(want['iterations'])
    bb, pp = np.nonzero(t.vrow >= 0)
    assert np.abs(V[t.vrow[bb, pp]].T - want['V'][:, bb, pp]).max() <= 1e-12
    ll, qq = np.nonzero(t.irow >= 0)
    assert np.abs(I[t.irow[ll, qq]].T - want['I'][:, ll, qq]).max() <= 1e-12


def test_dead_parent_stores_are_exactly_the_overwritten_ones():
    """Of the children that rewrite one phase of a bus only the earliest in topo order keeps its store."""
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.dss_reader import read_dss
    from gigpower_b200.flatten import flatten_fbs
    t = flatten_fbs(Circuit(read_dss(AUTHORED)))
    _, kind = _schedule(t)
    writers = {}
    for i in range(t.n_steps):
        for p in range(3):
            if t.bus_vrow[i, p] >= 0 and t.par_vrow[i, p] >= 0:
                writers.setdefault(int(t.par_vrow[i, p]), []).append((i, p))
    n_dead = 0
    for row, w in writers.items():
        for j, (i, p) in enumerate(w):
            assert bool((kind[i] >> (19 + p)) & 1) == (j > 0)
            n_dead += j > 0
    assert n_dead > 50          # IEEE 123: a large share of the backward sweep's V stores

"""The tree-walk FBS kernel (fbs_walk_kernel) checked WITHOUT a GPU: ``gp_fbs_walk_events`` (host only)
returns the event table the kernel is given; this test runs the kernel's loop over exactly that table in
NumPy - the current bus's voltage and child-current sum in "registers" (arrays cv, ca), one operand
request ahead, parked partial sums in the bus's own I rows, a single V array - and holds V, I and the
iteration counts to the oracle (the reference's two full sweeps).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import ALL_FEEDERS, ZIPS, load_model, REPO

DESC, LEAF, ASC = 0, 1, 2


def _events(t):
    from gigpower_b200 import _cabi
    d, keep = _cabi.make_fbs_desc(t)
    buf = np.zeros((2 * max(t.n_steps, 1), 20), dtype=np.int32)
    n = _cabi.gp_fbs_walk_events(C.byref(d), buf.ctypes.data_as(_cabi.i32p), len(buf))
    assert n >= 0
    return buf[:n]


def emulate_walk(t, rows, ev, maxiter=100):
    S = rows.shape[1]
    V = np.full((t.n_vrows, S), np.nan + 0j)                 # every row must be written before it is read
    I = np.full((max(t.n_irows, 1), S), np.nan + 0j)
    z = np.asarray(t.z).reshape(t.n_steps, 3, 3, 2)
    Z = z[..., 0] + 1j * z[..., 1]
    vref = np.asarray(t.vref).reshape(3, 2)
    vref = vref[:, 0] + 1j * vref[:, 1]
    aPQ, aI, aZ = [float(x) for x in t.zip]
    gam = np.asarray(t.vr_gamma).reshape(t.n_steps, 3)
    cap = np.asarray(t.cappu).reshape(t.n_steps, 3)
    active = np.ones(S, dtype=bool)
    its = np.zeros(S, dtype=int)

    def request(e, a, first):
        """operands of event e requested ahead: oI (edge currents or the parked sum), oS (loads)"""
        typ = e[1] & 3
        oI = np.zeros((3, a.sum()), dtype=complex)
        oS = np.zeros((3, a.sum()), dtype=complex)
        for p in range(3):
            if e[5 + p] >= 0 and not (first and typ != ASC):
                oI[p] = I[e[5 + p], a]
            if e[14 + p] >= 0:
                oS[p] = rows[e[14 + p], a]
        return oI, oS

    for _ in range(maxiter):
        if not active.any():
            break
        a = active.copy()
        first = _ == 0
        n = int(a.sum())
        cv = np.tile(vref[:, None], (1, n)).astype(complex)
        ca = np.zeros((3, n), dtype=complex)
        for p in range(3):
            V[t.root_vrow[p], a] = vref[p]
        nxt = request(ev[0], a, first)
        for j, e in enumerate(ev):
            oI, oS = nxt
            if j + 1 < len(ev):
                nxt = request(ev[j + 1], a, first)             # before this event's stores, as in the kernel
            i, fl = int(e[0]), int(e[1])
            typ, vr = fl & 3, bool(fl & 4)
            mb, ms, mw, mc, ml, me = [(fl >> s) & 7 for s in (3, 6, 9, 12, 15, 18)]
            e_row, ld_v, st_ca = e[2:5], e[8:11], e[11:14]
            brow, prow = t.bus_vrow[i], t.par_vrow[i]
            if typ == DESC:
                for p in range(3):
                    if st_ca[p] >= 0:
                        I[st_ca[p], a] = ca[p]
            rV = np.zeros((3, n), dtype=complex)
            if typ == ASC:
                for p in range(3):
                    if ld_v[p] >= 0:
                        rV[p] = V[ld_v[p], a]
            if vr:
                b = cv.copy()
                if typ != ASC:
                    for p in range(3):
                        b[p] = gam[i, p] * cv[p] if (mb >> p) & 1 else 0
                        if (mb >> p) & 1:
                            V[brow[p], a] = b[p]
                if typ == DESC:
                    cv, ca = b, np.zeros((3, n), dtype=complex)
                    continue
                for p in range(3):
                    nv = 0
                    if (mw >> p) & 1:
                        nv = (1.0 / gam[i, p]) * b[p]
                        V[prow[p], a] = nv
                    if typ == ASC:
                        cv[p] = nv if (mw >> p) & 1 else rV[p]
                        ca[p] = oI[p]
                    elif (mw >> p) & 1:
                        cv[p] = nv
                continue
            if typ != ASC:
                b = np.zeros((3, n), dtype=complex)
                for p in range(3):
                    if (mb >> p) & 1:
                        v = cv[p].copy()
                        if not first:
                            for q in range(3):
                                if (me >> q) & 1:
                                    v = v - Z[i, p, q] * oI[q]
                        V[brow[p], a] = v
                        b[p] = v
                if typ == DESC:
                    cv, ca = b, np.zeros((3, n), dtype=complex)
                    continue
            else:
                b = cv.copy()
            In = np.zeros((3, n), dtype=complex)
            for p in range(3):
                if (me >> p) & 1:
                    if typ == ASC:
                        In[p] = ca[p]
                    if (ml >> p) & 1:
                        mag = np.abs(b[p])
                        sv = oS[p] * (aPQ + aI * mag + aZ * mag ** 2) - 1j * cap[i, p] * mag ** 2
                        with np.errstate(all='ignore'):
                            In[p] = In[p] + np.where(b[p] != 0, np.conj(sv / np.where(b[p] != 0, b[p], 1)), 0)
                    else:
                        with np.errstate(all='ignore'):
                            tt = (b[p].real - b[p].real) + (b[p].imag - b[p].imag)
                        In[p] = In[p] + tt + 1j * tt
            In = np.nan_to_num(In)
            for p in range(3):
                if (ms >> p) & 1:
                    I[e_row[p], a] = In[p]
            for p in range(3):
                nv = 0
                if (mw >> p) & 1:
                    nv = b[p].copy()
                    for q in range(3):
                        if (me >> q) & 1:
                            nv = nv + Z[i, p, q] * In[q]
                    V[prow[p], a] = nv
                if typ == ASC:
                    cv[p] = nv if (mw >> p) & 1 else rV[p]
                    ca[p] = oI[p]
                elif (mw >> p) & 1:
                    cv[p] = nv
                if (mc >> p) & 1:
                    ca[p] = ca[p] + In[p]
        its[a] += 1
        diff = np.max(np.abs(cv - vref[:, None]), axis=0)
        done = np.zeros(S, dtype=bool)
        done[a] = diff <= t.tolerance
        active &= ~done
    return V, I, its


AUTHORED = os.path.join(REPO, 'gigpower_b200', 'feeders', 'ieee123_allwye_noxfm_noreg.dss')


def _tables(feeder):
    from gigpower_b200.circuit import Circuit
    from gigpower_b200.dss_reader import read_dss
    from gigpower_b200.flatten import flatten_fbs
    from gigpower_b200.solution import Solution
    Solution.set_zip_values(np.asarray(ZIPS['test_zip']))
    model = read_dss(AUTHORED) if feeder == 'ieee123' else load_model(feeder)
    return model, flatten_fbs(Circuit(model))


@pytest.mark.parametrize('feeder', ALL_FEEDERS + ['ieee123'])
def test_walk_loop_over_the_library_event_table_matches_the_oracle(feeder):
    from oracle.gp_oracle import FBSOracle
    model, t = _tables(feeder)
    ev = _events(t)
    if len(ev) == 0:
        pytest.skip('feeder has V rows below a regulator that no forward step writes: two-sweep kernel')
    steps = ev[:, 0]
    types = ev[:, 1] & 3
    assert sorted(steps[types != ASC]) == list(range(t.n_steps))      # every bus goes down exactly once
    assert sorted(steps[types == ASC]) == sorted(steps[types == DESC])  # and comes up if it has children
    o = FBSOracle(model, ZIPS['test_zip'])
    rng = np.random.Generator(np.random.PCG64(6))
    S = 24
    mult = rng.uniform(0.4, 1.5, size=(S, len(o.c.load_names)))
    mult[0] = 1.0
    spu = o.c.spu_matrix(o.c.load_kw * mult, o.c.load_kvar * mult)
    want = o.solve(spu)
    b, p = np.nonzero(t.srow >= 0)
    rows = np.zeros((max(t.n_srows, 1), S), dtype=complex)
    rows[t.srow[b, p]] = spu[:, b, p].T
    V, I, its = emulate_walk(t, rows, ev)
    assert list(its) == list(want['iterations'])
    bb, pp = np.nonzero(t.vrow >= 0)
    assert np.abs(V[t.vrow[bb, pp]].T - want['V'][:, bb, pp]).max() <= 1e-12
    # I rows the kernel stores itself (shared rows are filled in by the copy-out at the end of the kernel)
    stored = set()
    for e in ev:
        for q in range(3):
            if (e[1] >> (6 + q)) & 1:
                stored.add(int(e[2 + q]))
    ll, qq = np.nonzero(t.irow >= 0)
    keep = np.array([int(t.irow[l, q]) in stored for l, q in zip(ll, qq)])
    assert keep.sum() >= t.n_irows // 3
    assert np.abs(I[t.irow[ll, qq][keep]].T - want['I'][:, ll[keep], qq[keep]]).max() <= 1e-12


def test_walk_table_of_the_123_bus_feeder_has_the_expected_shape():
    model, t = _tables('ieee123')
    ev = _events(t)
    types = ev[:, 1] & 3
    assert (types == LEAF).sum() == 41 and (types == DESC).sum() == 84 and (types == ASC).sum() == 84
    # parked partial sums only at branch points
    assert 0 < (ev[:, 11:14] >= 0).any(axis=1).sum() <= 40

"""CPU oracle for gigpower's FBS and NR3 power flow -- TEST INFRASTRUCTURE.

A plain-NumPy restatement of the reference algorithms on flat arrays, batched
over scenarios.  It is the checker for the CUDA path and the ``cpu_baseline``
of ``bench.py``; it is NEVER imported by the product package
(``gigpower_b200``), which fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the UNMODIFIED reference
(``/root/reference/gigpower/src``) behind ``oracle/standin/opendssdirect`` in
the build container and commits its full-precision outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` holds this file to those
vectors (<= 1e-12, equal iteration counts) and to the reference's own recorded
2-decimal tables (``gigpower/tests/test_compare_{fbs,nr3}_dss/*.out.txt``).

Every function cites the reference code it follows (paths relative to
``/root/reference/gigpower/src/gigpower``).
"""
from __future__ import annotations

import numpy as np

SBASE = 10 ** 6          # circuit_element.py:25, circuit.py:36
MAXITER = 100            # solution.py:26
DEFAULT_ZIP = (0.10, 0.05, 0.85, 0.10, 0.05, 0.85, 0.80)   # solution.py:56-60


# ==========================================================================
# data model (circuit.py + element/group classes), flat
# ==========================================================================
def _bus_of(spec):                      # utils.py:6-13
    return spec.split('.')[0]


def _spec_phases(spec):                 # utils.py:16-27
    return [int(p) for p in spec.split('.')[1:]] if '.' in spec else [1, 2, 3]


def _phase_matrix(phases):              # utils.py:30-47
    m = np.zeros(3, dtype=int)
    for p in phases:
        m[int(p) - 1] = 1
    return m


def _pad_phases(mat, phases01):         # utils.py:74-102 (2-D branch)
    out = np.zeros((3, 3), dtype=mat.dtype)
    vals = iter(mat.flatten())
    for r in range(3):
        for c in range(3):
            if phases01[r] == 1 and phases01[c] == 1:
                try:
                    out[r][c] = next(vals)
                except StopIteration:
                    pass
    return out


class OracleCircuit:
    """Flat restatement of ``Circuit(dss)`` (circuit.py:36-64) for one feeder
    model (a ``gigpower_b200.dss_reader.FeederModel`` or any object with the
    same fields) and one ZIP vector."""

    def __init__(self, model, zip_v=DEFAULT_ZIP):
        zip_v = np.asarray(zip_v, dtype=float)
        self.aZ_p, self.aI_p, self.aPQ_p = zip_v[0:3]          # circuit.py:31-33
        self.bus_names = list(model.bus_names)
        nb = self.nb = len(self.bus_names)
        bidx = self.bidx = {n: i for i, n in enumerate(self.bus_names)}
        # Bus: circuit_element.py:40-51
        self.bus_phases = [list(model.bus_nodes[n]) for n in self.bus_names]
        self.bus_pm = np.array([_phase_matrix(p) for p in self.bus_phases])
        vbase = np.array([model.bus_kvbase[n] * 1000 for n in self.bus_names])
        ibase = SBASE / vbase
        self.zbase = vbase / ibase

        # Line: line.py:10-73
        self.line_names = [ln.name for ln in model.lines]
        nl = self.nl = len(model.lines)
        self.line_tx = np.array([bidx[_bus_of(ln.bus1)] for ln in model.lines], dtype=int)
        self.line_rx = np.array([bidx[_bus_of(ln.bus2)] for ln in model.lines], dtype=int)
        self.line_pm = np.zeros((nl, 3), dtype=int)
        self.FZpu = np.zeros((nl, 3, 3), dtype=complex)
        self.rmat = np.zeros((nl, 9))
        self.xmat = np.zeros((nl, 9))
        for i, ln in enumerate(model.lines):
            phases = [1, 2, 3] if ln.is_switch else _spec_phases(ln.bus1)   # line.py:28-36
            pm = _phase_matrix(phases)
            self.line_pm[i] = pm
            zb = self.zbase[self.line_tx[i]]
            fz_mult = 1 / zb * ln.length                                      # line.py:39
            RM = np.asarray(ln.rmatrix, dtype=float)
            XM = np.asarray(ln.xmatrix, dtype=float)
            ZM = fz_mult * (RM + 1j * XM)
            nph = len(phases)
            ZM = np.reshape(ZM, (ZM.shape[0] // nph, nph))
            self.FZpu[i] = _pad_phases(ZM, pm)                                # line.py:46
            # line.py:48-73
            if len(XM) == 1:
                xm, rm = (np.identity(3) * XM).flatten(), (np.identity(3) * RM).flatten()
            elif len(XM) == 9:
                xm, rm = XM, RM
            elif len(XM) == 4:
                xm = _pad_phases(np.reshape(XM, (2, 2)), pm).flatten()
                rm = _pad_phases(np.reshape(RM, (2, 2)), pm).flatten()
            else:
                raise IndexError(f"Xmat is length {len(XM)}, expected 1, 4, or 9 elements")
            self.xmat[i] = xm * (1 / zb * ln.length)
            self.rmat[i] = rm * (1 / zb * ln.length)

        # Load: load.py:9-35 ; LoadGroup: load_group.py:17-52
        self.load_names = [ld.name for ld in model.loads]
        self.load_bus = np.array([bidx[_bus_of(ld.bus1)] for ld in model.loads], dtype=int)
        self.load_pm = np.array([_phase_matrix(_spec_phases(ld.bus1)) for ld in model.loads]
                                ).reshape(-1, 3)
        self.load_nph = np.array([len(_spec_phases(ld.bus1)) for ld in model.loads], dtype=int)
        self.load_kw = np.array([ld.kw for ld in model.loads], dtype=float)
        self.load_kvar = np.array([ld.kvar for ld in model.loads], dtype=float)
        self.zip_mask = np.zeros((nb, 3))          # 1 where some load touches the bus-phase
        for j in range(len(model.loads)):
            self.zip_mask[self.load_bus[j], self.load_pm[j] == 1] = 1
        self.aPQ = self.zip_mask * self.aPQ_p      # load_group.py:39-52, circuit.py:140-162
        self.aI = self.zip_mask * self.aI_p
        self.aZ = self.zip_mask * self.aZ_p

        # Capacitor: capacitor.py:7-10 (phases = ALL nodes of its bus, circuit_element.py:47-51)
        self.cappu = np.zeros((nb, 3))
        for cp in model.capacitors:
            b = bidx[_bus_of(cp.bus1)]
            nph = len(self.bus_phases[b])
            self.cappu[b] += self.bus_pm[b] * (cp.kvar * 1000 / SBASE / nph)
        self.wpu = np.zeros((nb, 3))               # circuit.py:164-171

        # Transformer (not regulators): transformer_group.py:10-20, transformer.py:8-37
        regs = [rc.name for rc in model.regcontrols]
        xf = [t for t in model.transformers if t.name not in regs]
        self.xfm_names = [t.name for t in xf]
        self.ntf = len(xf)
        self.xfm_tx = np.array([bidx[_bus_of(t.buses[0])] for t in xf], dtype=int)
        self.xfm_rx = np.array([bidx[_bus_of(t.buses[1])] for t in xf], dtype=int)
        self.xfm_pm = np.array([_phase_matrix(_spec_phases(t.buses[0])) for t in xf]).reshape(-1, 3)
        self.xfm_nph = [len(_spec_phases(t.buses[0])) for t in xf]

        # VoltageRegulator: voltage_regulator.py:10-62
        self.vr_names = list(regs)
        self.nvr = len(regs)
        tmap = {t.name: t for t in model.transformers}
        self.vr_tx, self.vr_reg, self.vr_phases, self.vr_gamma = [], [], [], []
        for rc in model.regcontrols:
            t = tmap[rc.transformer]
            self.vr_tx.append(bidx[_bus_of(t.buses[0])])
            self.vr_reg.append(bidx[_bus_of(t.buses[1])])
            self.vr_phases.append(_spec_phases(t.buses[0]))
            g = (rc.tap + 16) / 32                   # voltage_regulator.py:23-39
            g *= 1.10 - .90
            g += .90
            self.vr_gamma.append(g)
        self.vr_pm = np.array([_phase_matrix(p) for p in self.vr_phases]).reshape(-1, 3)

    # load.py:16-24 + load_group.py:17-37 for per-load kW/kvar (shape [..., n_loads])
    def spu_matrix(self, kw=None, kvar=None):
        """``LoadGroup.get_spu_matrix`` for given per-load kW/kvar; leading
        batch dimensions allowed.  Returns ``[..., nb, 3]`` complex."""
        kw = self.load_kw if kw is None else np.asarray(kw, dtype=float)
        kvar = self.load_kvar if kvar is None else np.asarray(kvar, dtype=float)
        lead = kw.shape[:-1]
        spu = np.zeros(lead + (self.nb, 3), dtype=complex)
        for j in range(len(self.load_names)):
            ppu = kw[..., j] / 1000 / self.load_nph[j]
            qpu = kvar[..., j] / 1000 / self.load_nph[j]
            s = (self.load_pm[j] * ppu[..., None]) + 1j * (self.load_pm[j] * qpu[..., None])
            spu[..., self.load_bus[j], :] += s
        return spu


# ==========================================================================
# FBS  (solution_fbs.py)
# ==========================================================================
def _fbs_adjacency(c: OracleCircuit):
    """``SolutionFBS.set_adj`` (solution_fbs.py:39-74) on bus indices."""
    adj = {}
    for i in range(c.nl):                                   # line_group.py:62-72
        adj.setdefault(c.line_tx[i], []).append(c.line_rx[i])
    for i in range(c.ntf):
        adj.setdefault(c.xfm_tx[i], []).append(c.xfm_rx[i])
    vr_adj = {}
    for i in range(c.nvr):                                  # voltage_regulator_group.py:25-35
        vr_adj.setdefault(c.vr_tx[i], []).append(c.vr_reg[i])
        vr_adj.setdefault(c.vr_reg[i], []).append(c.vr_tx[i])
    for node, nbrs in vr_adj.items():                       # get_adj_set :60-72 (de-dup)
        uniq = []
        for v in nbrs:
            if v not in uniq:
                uniq.append(v)
        adj.setdefault(node, [])
        adj[node] = adj[node] + uniq
    for i in range(c.nvr):                                  # remove upstream VR legs
        try:
            adj[c.vr_reg[i]].remove(c.vr_tx[i])
        except (KeyError, ValueError):
            pass
    return adj


def _topo_sort(nb, adj):
    """``topo_sort``/``topo_sort_dfs`` (utils.py:130-167): reverse DFS finishing
    order, DFS roots in bus-index order, children in adjacency order."""
    order = [None] * nb
    status = ['new'] * nb
    clock = [nb]

    def dfs(u):
        status[u] = 'active'
        for v in adj.get(u, []):
            if status[v] == 'new':
                dfs(v)
            elif status[v] == 'active':
                raise ValueError('Not radial. Contains a cycle.')
        status[u] = 'finished'
        order[clock[0] - 1] = u
        clock[0] -= 1

    import sys
    sys.setrecursionlimit(max(10000, 4 * nb))
    for u in range(nb):
        if status[u] == 'new':
            dfs(u)
    return order


class FBSOracle:
    """Batched restatement of ``SolutionFBS`` (solution_fbs.py:27-136)."""

    def __init__(self, model, zip_v=DEFAULT_ZIP):
        c = self.c = OracleCircuit(model, zip_v)
        self.adj = _fbs_adjacency(c)
        self.topo_order = _topo_sort(c.nb, self.adj)
        self.root = self.topo_order[0]
        self.parents = {}
        for u, ch in self.adj.items():                       # utils.py:170-181
            for v in ch:
                self.parents.setdefault(v, []).append(u)
        self.Vref = np.array([1, np.exp(1j * 240 * np.pi / 180),
                              np.exp(1j * 120 * np.pi / 180)], dtype=complex)   # :34-35
        self.tolerance = abs((self.Vref[1]) * 10 ** -6)                           # :37
        # edge lookup (line_group.py:26-38): lines first, then transformers, then VRs
        self.edge = {}
        for i in range(c.nvr):
            key = (c.vr_tx[i], c.vr_reg[i])
            self.edge.setdefault(key, ('vr', []))[1].append(i)
        for i in range(c.ntf):
            self.edge[(c.xfm_tx[i], c.xfm_rx[i])] = ('line', c.nl + i, np.zeros((3, 3), complex),
                                                     c.xfm_pm[i])
        for i in range(c.nl):
            self.edge[(c.line_tx[i], c.line_rx[i])] = ('line', i, c.FZpu[i], c.line_pm[i])
        self.nrows_I = c.nl + c.ntf + c.nvr                                       # circuit.py:173-184

    # solution.py:177-220
    def _calc_sV(self, V, spu, b=None, wpu=None, droop=None):
        c = self.c
        w = c.wpu if wpu is None else wpu          # Solution.wpu (solution.py:106, 188-203): [nb,3] or [S,nb,3]
        if droop is not None:
            # volt-var controllers that set wpu = get_Q(|V|) every time calc_sV is evaluated (the hook of
            # oracle/make_golden_vvc_iter.py around the reference's calc_sV): droop = (ctl, points, minQ, maxQ)
            ctl, points, minQ, maxQ = droop
            w = np.broadcast_to(w, V.shape).copy() if np.ndim(w) == 2 else np.array(w, dtype=float)
            for k, p in ctl:
                if b is None or k == b:
                    w[:, k, p] = vvc_get_Q(np.abs(V[:, k, p]), points, minQ, maxQ)
        if b is None:
            aV = np.abs(V)
            upd = spu * (c.aPQ + c.aI * aV + c.aZ * aV ** 2) - 1j * c.cappu * aV ** 2 + 1j * w
            upd[..., c.bus_pm == 0] = 0
            return upd
        aV = np.abs(V[:, b])
        upd = spu[:, b] * (c.aPQ[b] + c.aI[b] * aV + c.aZ[b] * aV ** 2) \
            - 1j * c.cappu[b] * aV ** 2 + 1j * (w[b] if w.ndim == 2 else w[:, b])
        upd[:, c.bus_pm[b] == 0] = 0
        return upd

    def _sweep(self, V, I, spu, wpu=None, droop=None):
        """One pass of the while-body of ``solve`` (:88-121) on arrays
        ``V[S,nb,3]``, ``I[S,rows,3]`` in place."""
        c = self.c
        V[:, self.root] = self.Vref                                               # :89
        for bus in self.topo_order:                                               # :92-98
            for child in self.adj.get(bus, []):
                e = self.edge[(bus, child)]
                if e[0] == 'vr':                                                  # vr_forward :138-160
                    for k in e[1]:
                        ph = np.array(c.vr_phases[k]) - 1
                        V[:, child, ph] = c.vr_gamma[k] * V[:, bus, ph]
                else:                                                             # :162-184
                    _, row, Z, _pm = e
                    newV = V[:, bus] - np.einsum('ij,sj->si', Z, I[:, row])
                    V[:, child] = newV * c.bus_pm[child]
        sV = self._calc_sV(V, spu, wpu=wpu, droop=droop)                          # :101
        for bus in reversed(self.topo_order):                                     # :104-121
            parents = self.parents.get(bus, [])
            if len(parents) > 1:
                raise ValueError(f"Network not radial. Bus {bus} has parents {parents}")
            if len(parents) == 1:
                parent = parents[0]
                sV[:, bus] = self._calc_sV(V, spu, bus, wpu=wpu, droop=droop)     # :114
                e = self.edge[(parent, bus)]
                if e[0] == 'vr':
                    # vr_current :234-254 -> Itx stays 0 (aliased row never written)
                    for k in e[1]:                                                # vr_backward :186-205
                        ph = np.array(c.vr_phases[k]) - 1
                        V[:, parent, ph] = 1 / c.vr_gamma[k] * V[:, bus, ph]
                    continue
                _, row, Z, pm = e
                # update_current :256-288
                Vb = V[:, bus]
                newI = np.zeros_like(Vb)
                nz = Vb != 0
                with np.errstate(all='ignore'):
                    newI[nz] = np.conj(sV[:, bus][nz] / Vb[nz])
                for child in self.adj.get(bus, []):
                    ce = self.edge[(bus, child)]
                    if ce[0] == 'line':
                        newI = newI + I[:, ce[1]]
                    # VR child: += vr.Itx == 0
                I[:, row] = np.nan_to_num(newI * pm)
                # update_voltage_backward :207-232
                ph = np.array(c.bus_phases[bus]) - 1
                V[:, parent, ph] = V[:, bus, ph] + np.einsum('ij,sj->si', Z[ph], I[:, row])
                V[:, parent] = V[:, parent] * c.bus_pm[parent]
        return V, I

    def solve(self, spu=None, maxiter=MAXITER, V0=None, I0=None, wpu=None, droop=None):
        """Flat-start solve of ``S`` scenarios; ``spu[S,nb,3]`` complex (default:
        the feeder's nominal loads, ``S=1``).  Each scenario stops at its own
        iteration exactly as a separate reference object would.

        ``V0, I0``: warm start - what a reference object does whose ``Vtest`` and
        ``iterations`` are reset before ``solve()`` is called again: ``self.V`` and
        ``self.I`` keep the earlier solution (:76-90 never re-initialise them).
        ``wpu[S,nb,3]`` real: per-scenario ``Solution.wpu`` (the controllable reactive
        injection of ``calc_sV``, solution.py:106, 203; all zeros unless a caller sets it).
        ``droop = (ctl, points, minQ, maxQ)``: volt-var controllers at the (bus, phase) pairs ``ctl`` whose
        injection is ``wpu = VoltVARController.get_Q(|V|)`` (volt_var_controller.py:20-38) re-evaluated from the
        current voltage at every ``calc_sV``, i.e. inside every sweep."""
        c = self.c
        if spu is None:
            spu = c.spu_matrix()[None]
        spu = np.asarray(spu, dtype=complex)
        S = spu.shape[0]
        V = np.zeros((S, c.nb, 3), dtype=complex)                                 # solution.py:108-125
        I = np.zeros((S, self.nrows_I, 3), dtype=complex)
        if V0 is not None:
            V[:], I[:] = V0, I0
        iters = np.zeros(S, dtype=np.int32)
        diff = np.full(S, -1.0)
        Vtest = np.zeros((S, 3), dtype=complex)
        converged = np.max(np.abs(Vtest - self.Vref), axis=1) <= self.tolerance   # :85
        active = (~converged) & (iters < maxiter)
        while active.any():                                                       # :87
            idx = np.nonzero(active)[0]
            Vs, Is = self._sweep(V[idx], I[idx], spu[idx], None if wpu is None else wpu[idx], droop)
            V[idx], I[idx] = Vs, Is
            iters[idx] += 1                                                       # :123
            Vtest[idx] = Vs[:, self.root]                                         # :125
            with np.errstate(all='ignore'):
                diff[idx] = np.max(np.abs(Vtest[idx] - self.Vref), axis=1)        # :126
                converged[idx] = diff[idx] <= self.tolerance                      # :128
            active = (~converged) & (iters < maxiter)
        sV = self._calc_sV(V, spu, wpu=wpu, droop=droop)                          # :132
        Vmag = np.abs(V)                                                          # solution.py:232
        nl = c.nl
        Srx = V[:, c.line_rx] * np.conj(I[:, :nl])                                # solution.py:227-230
        Stx = V[:, c.line_tx] * np.conj(I[:, :nl])                                # solution.py:222-225
        return dict(V=V, I=I, sV=sV, Vmag=Vmag, Stx=Stx, Srx=Srx, iterations=iters,
                    converged=converged, convergence_diff=diff)


# ==========================================================================
# NR3  (solution_nr3.py + nr3_lib/)
# ==========================================================================
VSLACK = np.array([1, np.exp(1j * -120 * np.pi / 180), np.exp(1j * 120 * np.pi / 180)],
                  dtype=complex)                                                  # solution.py:23-24


class NR3Oracle:
    """Batched restatement of ``SolutionNR3`` (solution_nr3.py:28-663).

    The reference stores each KCL row as a dense ``N x N`` matrix ``H[i]``; the
    same rows are kept here as ``{(a, b): value}`` maps with the reference's
    ASSIGNMENT semantics (a later write to the same entry overwrites), and are
    evaluated as ``X' H X`` / ``2 X' H``.
    """

    def __init__(self, model, zip_v=DEFAULT_ZIP):
        c = self.c = OracleCircuit(model, zip_v)
        nb, nl = c.nb, c.nl
        self.tf_lines = int(sum(c.xfm_nph))                   # line_group.py:100-102
        self.vr_lines = int(sum(len(p) for p in c.vr_phases))
        self.N = 2 * 3 * (nb + nl) + 2 * self.tf_lines + 2 * 2 * self.vr_lines   # :53
        self._slack()
        self._kvl()
        self._kcl_structure()
        self._kvl_vregs()

    # index helpers (solution_nr3.py:58-61, 141-161)
    def vA(self, ph, k):
        return 2 * self.c.nb * ph + 2 * k

    def iC(self, ph, l):
        return 2 * 3 * self.c.nb + 2 * ph * self.c.nl + 2 * l

    def _slack(self):                                          # :84-109
        nb = self.c.nb
        self.g_SB = np.zeros((6, self.N))
        sb_idx = [0, 1, 2 * nb, 2 * nb + 1, 4 * nb, 4 * nb + 1]
        for i in range(6):
            self.g_SB[i, sb_idx[i]] = 1
        self.b_SB = np.zeros(6)
        for i in range(3):
            self.b_SB[2 * i] = VSLACK[i].real
            self.b_SB[2 * i + 1] = VSLACK[i].imag

    def _kvl(self):                                            # :111-177
        c, nb, nl = self.c, self.c.nb, self.c.nl
        rows = 2 * 3 * nl + 2 * self.tf_lines + 2 * 2 * self.vr_lines
        G = np.zeros((rows, self.N))
        R, X = c.rmat, c.xmat
        for ph in range(3):
            for line in range(nl):
                b1, b2 = c.line_tx[line], c.line_rx[line]
                lp = c.line_pm[line]
                r0 = 2 * ph * nl + 2 * line
                if lp[ph] == 1:
                    G[r0][2 * nb * ph + 2 * b1] = 1
                    G[r0][2 * nb * ph + 2 * b2] = -1
                    G[r0 + 1][2 * nb * ph + 2 * b1 + 1] = 1
                    G[r0 + 1][2 * nb * ph + 2 * b2 + 1] = -1
                    for k in range(3):
                        cc = 2 * 3 * nb + 2 * k * nl + 2 * line
                        G[r0][cc] = -R[line][ph * 3 + k] * lp[k]
                        G[r0][cc + 1] = X[line][ph * 3 + k] * lp[k]
                        G[r0 + 1][cc] = -X[line][ph * 3 + k] * lp[k]
                        G[r0 + 1][cc + 1] = -R[line][ph * 3 + k] * lp[k]
                else:
                    G[r0][2 * nb * 3 + 2 * ph * nl + 2 * line] = 1
                    G[r0 + 1][2 * nb * 3 + 2 * ph * nl + 2 * line + 1] = 1
        kvl_count = 0                                          # :163-175
        for t in range(c.ntf):
            for ph in range(3):
                if c.xfm_pm[t][ph] != 0:
                    line = kvl_count
                    r0 = 2 * 3 * nl + 2 * line
                    G[r0][2 * nb * ph + 2 * c.xfm_tx[t]] = 1
                    G[r0][2 * nb * ph + 2 * c.xfm_rx[t]] = -1
                    G[r0 + 1][2 * nb * ph + 2 * c.xfm_tx[t] + 1] = 1
                    G[r0 + 1][2 * nb * ph + 2 * c.xfm_rx[t] + 1] = -1
                    kvl_count += 1
        self.G_KVL = G
        self.b_KVL = np.zeros(rows)

    def _kcl_structure(self):                                  # :179-442, load-independent part
        c, nb, nl = self.c, self.c.nb, self.c.nl
        nrow = 2 * 3 * (nb - 1)
        H = [dict() for _ in range(nrow)]
        in_lines = [[] for _ in range(nb)]                     # line_group.py:172-188
        out_lines = [[] for _ in range(nb)]
        key2line = {}
        for l in range(nl):
            key2line[(c.line_tx[l], c.line_rx[l])] = l
        radj, adj = {}, {}
        for l in range(nl):
            adj.setdefault(c.line_tx[l], []).append(c.line_rx[l])
            radj.setdefault(c.line_rx[l], []).append(c.line_tx[l])
        for k in range(nb):
            in_lines[k] = [key2line[(tx, k)] for tx in radj.get(k, [])]
            out_lines[k] = [key2line[(k, rx)] for rx in adj.get(k, [])]
        self.diagA, self.diagB = {}, {}                        # row -> X index of A / B (when phase exists)
        for ph in range(3):
            for k2 in range(1, nb):
                avail = c.bus_pm[k2][ph] == 1
                for cplx in range(2):
                    row = 2 * ph * (nb - 1) + (k2 - 1) * 2 + cplx
                    a, b = 2 * nb * ph + 2 * k2, 2 * nb * ph + 2 * k2 + 1
                    if avail:
                        self.diagA[row], self.diagB[row] = a, b
                        H[row][(a, a)] = 0.0                   # filled per scenario (:256-261)
                        H[row][(b, b)] = 0.0
                    for sign, lst in ((1, in_lines[k2]), (-1, out_lines[k2])):      # :265-299
                        for l in lst:
                            if not avail:
                                continue
                            cc, dd = self.iC(ph, l), self.iC(ph, l) + 1
                            if cplx == 0:
                                H[row][(a, cc)] = H[row][(cc, a)] = sign * 0.5
                                H[row][(b, dd)] = H[row][(dd, b)] = sign * 0.5
                            else:
                                H[row][(a, dd)] = H[row][(dd, a)] = -sign * 0.5
                                H[row][(b, cc)] = H[row][(cc, b)] = sign * 0.5
        base_t = 2 * 3 * (nb + nl)

        def bilinear(k2, ph, cc, sign):
            row = 2 * ph * (nb - 1) + (k2 - 1) * 2
            a, b = 2 * nb * ph + 2 * k2, 2 * nb * ph + 2 * k2 + 1
            dd = cc + 1
            H[row][(a, cc)] = H[row][(cc, a)] = sign * 0.5
            H[row][(b, dd)] = H[row][(dd, b)] = sign * 0.5
            H[row + 1][(a, dd)] = H[row + 1][(dd, a)] = -sign * 0.5
            H[row + 1][(b, cc)] = H[row + 1][(cc, b)] = sign * 0.5

        # transformers :301-349
        for which, sign in ((1, 1), (0, -1)):
            count = 0
            for i in range(c.ntf):
                for ph in range(3):
                    k2 = int(c.xfm_rx[i] if which == 1 else c.xfm_tx[i])
                    if k2 == 0:
                        count += 1
                    if k2 != 0 and c.xfm_pm[i][ph] != 0:
                        bilinear(k2, ph, base_t + 2 * count, sign)
                        count += 1
        # voltage regulators :351-400
        base_r = base_t + 2 * self.tf_lines
        for which, sign, first in ((1, 1, 0), (0, -1, 1)):
            count = 0
            for i in range(c.nvr):
                for ph in range(3):
                    k2 = int(c.vr_reg[i] if which == 1 else c.vr_tx[i])
                    if k2 == 0:
                        count += 1
                    if k2 != 0 and c.vr_pm[i][ph] != 0:
                        line_idx = first + 2 * count          # line_in_idx_vr / line_out_idx_vr
                        bilinear(k2, ph, base_r + 2 * line_idx, sign)
                        count += 1
        self.H_struct = H
        # linear term g :414-419
        self.g = np.zeros((nrow, self.N))
        for ph in range(3):
            for k2 in range(1, nb):
                for cplx in range(2):
                    if c.bus_pm[k2][ph] == 0:
                        self.g[2 * (nb - 1) * ph + 2 * (k2 - 1) + cplx, 2 * ph * nb + 2 * k2 + cplx] = 1
        # flatten the bilinear structure into index arrays for vectorised evaluation
        ri, ai, bi, vi = [], [], [], []
        for r, d in enumerate(H):
            for (a, b), v in d.items():
                if a != b:
                    ri.append(r), ai.append(a), bi.append(b), vi.append(v)
        self._hr, self._ha, self._hb = (np.array(x, dtype=int) for x in (ri, ai, bi))
        self._hv = np.array(vi, dtype=float)

    def _kvl_vregs(self):                                      # :444-515
        c, nb, nl = self.c, self.c.nb, self.c.nl
        vrl = self.vr_lines
        self.G_reg = np.zeros((2 * vrl, self.N))
        Hreg = [dict() for _ in range(2 * vrl)]
        base_r = 2 * 3 * (nb + nl) + 2 * self.tf_lines
        cnt = 0
        for m in range(c.nvr):
            b1, b2 = c.vr_tx[m], c.vr_reg[m]
            gain = -1 * c.vr_gamma[m]                          # voltage_regulator_group.py:53-58
            for ph in range(3):
                if c.vr_pm[m][ph] != 0:
                    A1, B1 = 2 * nb * ph + 2 * b1, 2 * nb * ph + 2 * b1 + 1
                    A2, B2 = 2 * nb * ph + 2 * b2, 2 * nb * ph + 2 * b2 + 1
                    self.G_reg[2 * cnt][A1] = gain
                    self.G_reg[2 * cnt][A2] = 1
                    self.G_reg[2 * cnt + 1][B1] = gain
                    self.G_reg[2 * cnt + 1][B2] = 1
                    Co, Do = base_r + 2 * (2 * cnt + 1), base_r + 2 * (2 * cnt + 1) + 1
                    Ci, Di = base_r + 2 * (2 * cnt), base_r + 2 * (2 * cnt) + 1
                    h0, h1 = Hreg[2 * cnt], Hreg[2 * cnt + 1]
                    for (a, b, v) in ((A1, Co, 1), (B1, Do, 1), (A2, Ci, -1), (B2, Di, -1)):
                        h0[(a, b)] = h0[(b, a)] = v
                    for (a, b, v) in ((B1, Co, 1), (A1, Do, -1), (B2, Ci, -1), (A2, Di, 1)):
                        h1[(a, b)] = h1[(b, a)] = v
                    cnt += 1
        ri, ai, bi, vi = [], [], [], []
        for r, d in enumerate(Hreg):
            for (a, b), v in d.items():
                ri.append(r), ai.append(a), bi.append(b), vi.append(v)
        self._rr, self._ra, self._rb = (np.array(x, dtype=int) for x in (ri, ai, bi))
        self._rv = np.array(vi, dtype=float)

    # ---- per-scenario coefficients (:226-261, :403-440) ---------------------
    def _load_coeffs(self, p, q):
        """p, q: [S, nb, 3] per-bus ppu/qpu sums (load_group.py:25-37).  Returns
        hA, hB, b0 of shape [S, 6(nb-1)] in KCL row order."""
        c, nb = self.c, self.c.nb
        S = p.shape[0]
        nrow = 2 * 3 * (nb - 1)
        hA, hB, b0 = (np.zeros((S, nrow)) for _ in range(3))
        beta_S = gamma_S = c.aPQ_p
        beta_I = gamma_I = c.aI_p
        beta_Z = gamma_Z = c.aZ_p
        for ph in range(3):
            A0, B0 = (1, 0) if ph == 0 else ((-1 / 2, -1 * np.sqrt(3) / 2) if ph == 1
                                             else (-1 / 2, np.sqrt(3) / 2))
            h00 = -((A0 ** 2) * (A0 ** 2 + B0 ** 2) ** (-3 / 2)) + (A0 ** 2 + B0 ** 2) ** (-1 / 2)
            h11 = -((B0 ** 2) * (A0 ** 2 + B0 ** 2) ** (-3 / 2)) + ((A0 ** 2 + B0 ** 2) ** (-1 / 2))
            for k2 in range(1, nb):
                if c.bus_pm[k2][ph] != 1:
                    continue
                for cplx in range(2):
                    row = 2 * ph * (nb - 1) + (k2 - 1) * 2 + cplx
                    load = p[:, k2, ph] if cplx == 0 else q[:, k2, ph]
                    cap = 0 if cplx == 0 else c.cappu[k2][ph]
                    hA[:, row] = -load * (beta_Z + (0.5 * beta_I * h00)) + cap * (gamma_Z + (0.5 * gamma_I * h00))
                    hB[:, row] = -load * (beta_Z + (0.5 * beta_I * h11)) + cap * (gamma_Z + (0.5 * gamma_I * h11))
                    b_factor = 0 if cplx == 0 else c.cappu[k2][ph]
                    b0[:, row] = (-load * beta_S) + (b_factor * gamma_S)
        return hA, hB, b0

    # ---- residual / Jacobian (nr3_lib/compute_NR3FT.py, compute_NR3JT.py) ----
    def _FT_JT(self, X, hA, hB, b0):
        S = X.shape[0]
        nrow = hA.shape[1]
        FTSUBV = X @ self.g_SB.T - self.b_SB
        FTKVL = X @ self.G_KVL.T - self.b_KVL
        FTKCL = X @ self.g.T + b0
        JKCL = np.broadcast_to(self.g, (S,) + self.g.shape).copy()
        # off-diagonal bilinear entries: X' H X sums H[a,b] X[a] X[b]; 2 X' H puts 2 H[a,b] X[a] in col b
        np.add.at(FTKCL, (slice(None), self._hr), self._hv * X[:, self._ha] * X[:, self._hb])
        np.add.at(JKCL, (slice(None), self._hr, self._hb), 2 * self._hv * X[:, self._ha])
        rows = np.array(sorted(self.diagA), dtype=int)
        a = np.array([self.diagA[r] for r in rows], dtype=int)
        b = np.array([self.diagB[r] for r in rows], dtype=int)
        FTKCL[:, rows] += hA[:, rows] * X[:, a] ** 2 + hB[:, rows] * X[:, b] ** 2
        JKCL[:, rows, a] += 2 * hA[:, rows] * X[:, a]
        JKCL[:, rows, b] += 2 * hB[:, rows] * X[:, b]
        nvr2 = 2 * self.vr_lines
        FTVR = np.zeros((S, nvr2))
        JVR = np.zeros((S, nvr2, self.N))
        if nvr2:
            np.add.at(FTVR, (slice(None), self._rr), self._rv * X[:, self._ra] * X[:, self._rb])
            np.add.at(JVR, (slice(None), self._rr, self._rb), 2 * self._rv * X[:, self._ra])
        FTVR2 = X @ self.G_reg.T
        FT = np.concatenate([FTSUBV, FTKVL, FTKCL, FTVR, FTVR2], axis=1)
        JT = np.concatenate([np.broadcast_to(self.g_SB, (S,) + self.g_SB.shape),
                             np.broadcast_to(self.G_KVL, (S,) + self.G_KVL.shape),
                             JKCL, JVR,
                             np.broadcast_to(self.G_reg, (S,) + self.G_reg.shape)], axis=1)
        return FT, JT

    def x0(self, S):                                           # :37-82
        nb = self.c.nb
        X = np.zeros((S, self.N))
        for ph in range(3):
            for k in range(nb):
                X[:, 2 * ph * nb + 2 * k] = VSLACK[ph].real
                X[:, 2 * ph * nb + 2 * k + 1] = VSLACK[ph].imag
        return X

    def solve(self, spu=None, maxiter=MAXITER, chunk=16, return_history=False, X0=None):
        """Flat-start Newton solve of ``S`` scenarios (:592-632).  ``spu[S,nb,3]``
        holds the per-bus load sums (p = real, q = imag).  ``X0[S,N]``: start from an
        earlier solution instead (``solve()`` starts from ``self.XNR``, :598, which a
        second call finds at the previous solution)."""
        c = self.c
        if spu is None:
            spu = c.spu_matrix()[None]
        spu = np.asarray(spu, dtype=complex)
        S = spu.shape[0]
        X = self.x0(S) if X0 is None else np.array(X0, dtype=float)
        iters = np.zeros(S, dtype=np.int32)
        hist = [[] for _ in range(S)]
        for s0 in range(0, S, chunk):
            sl = slice(s0, min(S, s0 + chunk))
            hA, hB, b0 = self._load_coeffs(spu[sl].real, spu[sl].imag)
            Xc = X[sl]
            n = Xc.shape[0]
            FTmax = np.full(n, 1e99)
            it = np.zeros(n, dtype=np.int32)
            active = (FTmax >= 1e-9) & (it < maxiter)                             # :618
            while active.any():
                idx = np.nonzero(active)[0]
                FT, JT = self._FT_JT(Xc[idx], hA[idx], hB[idx], b0[idx])
                with np.errstate(all='ignore'):
                    FTmax[idx] = np.amax(np.abs(FT), axis=1)
                # (np.amax propagates NaN and ``NaN >= 1e-9`` is False: the reference's loop ends on the first residual
                # that holds a NaN - as nr3_solve_kernel's nanflag does - and goes on while it is merely infinite)
                if return_history:
                    for j, i in enumerate(idx):
                        hist[s0 + i].append(FTmax[i])
                if JT.shape[1] >= JT.shape[2]:                                    # :625-626
                    JtJ = np.swapaxes(JT, 1, 2) @ JT
                    try:
                        step = np.linalg.inv(JtJ) @ (np.swapaxes(JT, 1, 2) @ FT[:, :, None])
                    except np.linalg.LinAlgError:
                        step = np.full((len(idx), self.N, 1), np.nan)
                    Xc[idx] = Xc[idx] - step[:, :, 0]
                it[idx] += 1
                with np.errstate(invalid='ignore'):
                    active = (FTmax >= 1e-9) & (it < maxiter)
            X[sl] = Xc
            iters[sl] = it
        out = self.map_output(X, spu)
        out.update(XNR=X, iterations=iters)
        if return_history:
            out['history'] = hist
        return out

    @staticmethod
    def _scrub(a):
        """nr3_lib/map_output.py:24-27 etc.: ``|re| <= 1e-12 -> v = imag + 0j``
        (the imaginary part lands in the REAL slot), else ``|im| <= 1e-12 -> re``."""
        re, im = a.real.copy(), a.imag.copy()
        m1 = np.abs(re) <= 1e-12
        m2 = (~m1) & (np.abs(im) <= 1e-12)
        new_re = np.where(m1, im, re)
        new_im = np.where(m1 | m2, 0.0, im)
        return new_re + 1j * new_im

    def map_output(self, X, spu):                              # map_output.py:4-79, results [S, 3, n]
        c, nb, nl = self.c, self.c.nb, self.c.nl
        S = X.shape[0]
        V = np.zeros((S, 3, nb), dtype=complex)
        for ph in range(3):
            V[:, ph, :] = X[:, 2 * ph * nb:2 * ph * nb + 2 * nb:2] + 1j * X[:, 2 * ph * nb + 1:2 * ph * nb + 2 * nb:2]
        V = self._scrub(V)
        PH = c.bus_pm.T                                        # 3 x nb
        V[:, PH == 0] = 0
        Xl = X[:, 2 * 3 * nb:]
        I = np.zeros((S, 3, nl), dtype=complex)
        for ph in range(3):
            I[:, ph, :] = Xl[:, 2 * ph * nl:2 * ph * nl + 2 * nl:2] + 1j * Xl[:, 2 * ph * nl + 1:2 * ph * nl + 2 * nl:2]
        I = self._scrub(I)
        Stx = self._scrub(V[:, :, c.line_tx] * np.conj(I))
        Srx = self._scrub(V[:, :, c.line_rx] * np.conj(I))
        spuT = np.swapaxes(spu, 1, 2)                          # [S,3,nb]
        aV = np.abs(V)
        sV = spuT * (c.aPQ.T + c.aI.T * aV + c.aZ.T * aV ** 2) - 1j * c.cappu.T.real + 1j * c.wpu.T
        sV[:, PH == 0] = 0
        sV = self._scrub(sV)
        iN = np.zeros_like(sV)
        with np.errstate(all='ignore'):
            iN[:, PH != 0] = np.conj(sV[:, PH != 0] / V[:, PH != 0])
        iN = self._scrub(iN)
        return dict(V=V, I=I, Stx=Stx, Srx=Srx, i_Node=iN, sV=sV, Vmag=np.abs(V))


# ==========================================================================
# per-scenario regulator taps
# ==========================================================================
def gamma_of_tap(tap):
    """voltage_regulator.py:23-39, operation for operation."""
    g = (np.asarray(tap, dtype=float) + 16) / 32
    g = g * (1.10 - .90)
    g = g + .90
    return g


def solve_with_taps(oracle_cls, model, spu, taps, zip_v=DEFAULT_ZIP, **kw):
    """Scenario s = a fresh reference object whose regulators sit at ``taps[s]``
    (the reference reads a tap once, when the object is built:
    voltage_regulator.py:10-22).  ``taps[S, n_regcontrols]`` integers in the
    order of ``model.regcontrols``; scenarios are grouped by tap vector and each
    group is solved by an oracle built on a copy of ``model`` with those taps.
    Returns the oracle's result dict with the groups scattered back."""
    import copy
    spu = np.asarray(spu, dtype=complex)
    taps = np.asarray(taps, dtype=int).reshape(spu.shape[0], -1)
    out = {}
    uniq, inv = np.unique(taps, axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    for g, tv in enumerate(uniq):
        idx = np.nonzero(inv == g)[0]
        m = copy.deepcopy(model)
        for rc, t in zip(m.regcontrols, tv):
            rc.tap = int(t)
        r = oracle_cls(m, zip_v).solve(spu[idx], **kw)
        for k, v in r.items():
            if not isinstance(v, np.ndarray) or v.shape[:1] != (len(idx),):
                continue
            if k not in out:
                out[k] = np.zeros((spu.shape[0],) + v.shape[1:], dtype=v.dtype)
            out[k][idx] = v
    return out


# ==========================================================================
# volt-var droop around the FBS solve
# ==========================================================================
def vvc_get_Q(V_pu, points, minQ, maxQ):
    """``VoltVARController.get_Q`` (volt_var_controller.py:20-38) elementwise on an array."""
    V1, V2, V3, V4 = points
    V_pu = np.asarray(V_pu, dtype=float)
    out = np.empty_like(V_pu)
    for i, v in np.ndenumerate(V_pu):
        if v <= V1:
            q = maxQ
        elif v <= V2:
            slope = (- maxQ) / (V2 - V1)
            b = - (slope * V2)
            q = slope * v + b
        elif v <= V3:
            q = 0
        elif v <= V4:
            slope = (minQ) / (V4 - V3)
            b = - (slope * V3)
            q = slope * v + b
        else:
            q = minQ
        out[i] = -q
    return out


def fbs_volt_var(oracle: FBSOracle, spu, ctl, points, minQ, maxQ, rounds):
    """The quasi-static loop of ``oracle/make_golden_wpu.py`` on a batch: solve (round 0 from flat
    start, then from the last state), ``wpu[ctl] = get_Q(|V[ctl]|)``, again.  ``ctl``: (bus, phase)
    pairs.  Returns per round V ``[rounds,S,nb,3]``, Q ``[rounds,S,n_ctl]``, iterations ``[rounds,S]``."""
    spu = np.asarray(spu, dtype=complex)
    S = spu.shape[0]
    wpu = np.zeros((S, oracle.c.nb, 3))
    V0 = I0 = None
    Vs, Qs, its = [], [], []
    for _ in range(rounds):
        r = oracle.solve(spu, V0=V0, I0=I0, wpu=wpu)
        V0, I0 = r['V'], r['I']
        q = np.stack([vvc_get_Q(np.abs(r['V'][:, k, p]), points, minQ, maxQ) for k, p in ctl], axis=1)
        wpu = np.zeros((S, oracle.c.nb, 3))
        for j, (k, p) in enumerate(ctl):
            wpu[:, k, p] = q[:, j]
        Vs.append(r['V']), Qs.append(q), its.append(r['iterations'])
    return np.array(Vs), np.array(Qs), np.array(its)

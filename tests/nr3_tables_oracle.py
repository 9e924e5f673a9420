"""The tree-elimination Newton step of oracle/nr3_tree.py run over the PRODUCT's NR3 tables -- TEST INFRASTRUCTURE.

``gigpower_b200.flatten.flatten_nr3`` turns a Circuit into the ``Nr3Tables`` the CUDA kernels read (steps in sweep order,
packed V / I rows, 3x3 blocks, regulator ratios, capacitors, load rows, the map to the reference's XNR order).  On a box
without a GPU nothing exercised those tables: the tree oracle builds its own structures from the parsed feeder.  This
class fills the oracle's structures FROM the tables instead - topology from ``par_vrow`` / ``child_step``, impedances from
``z``, ratios from ``vr_gamma``, capacitors from ``cappu``, the ZIP constants from ``zip`` / ``cA`` / ``cB``, loads from
packed ``spu`` rows through ``load_srow`` - runs the oracle's own Newton step and maps the state to XNR through
``bus_vrow`` / ``edge_irow`` / ``edge_orow`` / ``xnr_index`` exactly as ``Nr3BatchResult.XNR`` does.  A wrong table entry
shows up as a different XNR than the reference's.  (EDGE_LINE = 0 covers lines and transformers, whose block is zero.)
"""
from types import SimpleNamespace

import numpy as np

from oracle.nr3_tree import NR3TreeOracle
from oracle.gp_oracle import MAXITER

EDGE_VR = 1


class NR3TablesOracle(NR3TreeOracle):
    def __init__(self, t):          # noqa: super().__init__ is replaced on purpose: everything comes from the tables
        self.t = t
        nb = t.vrow.shape[0]
        bus_of_vrow = {int(t.vrow[b, p]): b for b in range(nb) for p in range(3) if t.vrow[b, p] >= 0}
        bus_pm = (t.vrow >= 0).astype(int)
        cappu = np.zeros((nb, 3))
        cappu[0] = t.root_cappu
        self.order = [0]
        self.parent, self.edge_in, self.children = {}, {}, {k: [] for k in range(nb)}
        self.edges = []
        step_bus = []
        for i in range(t.n_steps):
            rows = t.bus_vrow[i]
            k = bus_of_vrow[int(rows[rows >= 0][0])]
            assert all(bus_of_vrow[int(r)] == k for r in rows[rows >= 0]) and np.array_equal(rows, t.vrow[k])
            prow = t.par_vrow[i]
            p = bus_of_vrow[int(prow[prow >= 0][0])]
            assert np.array_equal(prow, t.vrow[p])
            pm = (rows >= 0).astype(int)
            kind = 'vr' if (int(t.kind[i]) & 0xff) == EDGE_VR else 'line'
            Z = t.z[i, :, :, 0] + 1j * t.z[i, :, :, 1]
            self.edges.append(dict(kind=kind, tx=p, rx=k, Z=Z, pm=pm, gamma=np.asarray(t.vr_gamma[i], dtype=float)))
            self.order.append(k)
            self.parent[k], self.edge_in[k] = p, i
            cappu[k] = t.cappu[i]
            step_bus.append(k)
        assert self.order == [int(b) for b in t.order], 'steps are not in the order the tables announce'
        for i in range(t.n_steps):             # the child lists the backward sweep walks
            kids = [step_bus[int(s)] for s in t.child_step[t.child_begin[i]:t.child_begin[i] + t.child_count[i]]]
            assert all(self.parent[ch] == step_bus[i] for ch in kids)
            self.children[step_bus[i]] = kids
            for j, ch in enumerate(kids):      # a child's current leaves this bus through the child's OUT row
                assert np.array_equal(t.child_irow[t.child_begin[i] + j], t.edge_orow[self.edge_in[ch]])
        self.children[0] = [k for k in step_bus if self.parent[k] == 0]
        assert sum(len(v) for v in self.children.values()) == t.n_steps
        aPQ, aI, aZ = (float(x) for x in t.zip)
        self.c = SimpleNamespace(nb=nb, nl=t.n_lines, bus_pm=bus_pm, cappu=cappu, aPQ_p=aPQ, aI_p=aI, aZ_p=aZ)
        self.cA, self.cB, self.f0_absent, self.N = np.asarray(t.cA), np.asarray(t.cB), float(t.f0_absent), int(t.N)
        self.step_bus = step_bus

    def solve_rows(self, rows, maxiter=MAXITER):
        """rows: packed spu ``[n_srows, S]`` as ``solve_batch(spu_rows=...)`` takes them."""
        t = self.t
        S = rows.shape[1]
        spu = np.zeros((S, self.c.nb, 3), dtype=complex)
        for i, k in enumerate(self.step_bus):
            for p in range(3):
                if t.load_srow[i, p] >= 0:
                    spu[:, k, p] = rows[t.load_srow[i, p]]
        return self.solve(spu, maxiter=maxiter)

    def to_xnr(self, V, Iin, Iout):
        """State -> packed rows -> XNR through the tables' own maps (what Nr3BatchResult.XNR does on the device rows)."""
        t = self.t
        S = V.shape[0]
        X = np.zeros((t.n_vrows + t.n_irows, S), dtype=complex)
        for p in range(3):
            X[t.root_vrow[p]] = V[:, 0, p]
        for i, k in enumerate(self.step_bus):
            for p in range(3):
                if t.bus_vrow[i, p] >= 0:
                    X[t.bus_vrow[i, p]] = V[:, k, p]
                    X[t.n_vrows + t.edge_irow[i, p]] = Iin[:, i, p]
                    X[t.n_vrows + t.edge_orow[i, p]] = Iout[:, i, p] if self.edges[i]['kind'] == 'vr' else Iin[:, i, p]
        out = np.zeros((S, t.N))
        out[:, t.xnr_index] = X.real.T
        out[:, t.xnr_index + 1] = X.imag.T
        return out

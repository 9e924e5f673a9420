"""Tree-structured Newton step for NR3 -- TEST INFRASTRUCTURE (NumPy).

``NR3Oracle`` (gp_oracle.py) follows the reference literally: dense Jacobian,
``inv(J'J) J' F``.  That is O(N^3) and cannot be run on big feeders.  For a
radial feeder the same Newton step ``J dx = F`` can be solved exactly by block
elimination along the tree (a backward sweep that carries a 6x6 equivalent
admittance per bus, then a forward sweep) -- the algorithm the CUDA kernels in
``gigpower_b200/csrc/gp_nr3.cu`` implement.  This file is the NumPy statement
of that algorithm; ``tests/test_oracle_golden.py`` checks it against
``NR3Oracle`` and the reference's golden vectors, after which it serves as the
checker for feeders where the dense oracle is too slow.

Equations (reference ``src/solution_nr3.py``): per non-root bus k with
incoming line l from parent p, in terms of the Newton correction d(.):

  KVL_l (:141-158):  E dv_p - dv_k - Z di_l                  = r1
  KCL_k (:265-299):  D dv_k + M di_l - M sum_c di_c          = r2

with M = [[A, B], [B, -A]] per phase (d/dC, d/dD of the KCL rows), D the 2x2
d/dA, d/dB block, E the line-phase mask.  A child c contributes
di_c = a_c - Y_c dv_k, which gives (Dt = M^-1 D + sum Y_c, rt = M^-1 r2 + sum a_c)

  (I - Dt Z) di_l = rt + Dt r1 - Dt E dv_p   =>   di_l = a_k - Y_k dv_p
  dv_k = E dv_p - Z di_l - r1
"""
from __future__ import annotations

import numpy as np

from .gp_oracle import OracleCircuit, VSLACK, MAXITER, DEFAULT_ZIP, _fbs_adjacency, _topo_sort


def _real_rep(Z):
    """complex [3,3] -> real [6,6] acting on (re, im) interleaved vectors."""
    R = np.zeros((6, 6))
    R[0::2, 0::2] = Z.real
    R[0::2, 1::2] = -Z.imag
    R[1::2, 0::2] = Z.imag
    R[1::2, 1::2] = Z.real
    return R


class NR3TreeOracle:
    def __init__(self, model, zip_v=DEFAULT_ZIP):
        c = self.c = OracleCircuit(model, zip_v)
        adj = _fbs_adjacency(c)
        self.order = _topo_sort(c.nb, adj)
        if self.order[0] != 0:
            raise NotImplementedError('slack bus (index 0) must be the root')
        # edges: real lines, then transformers (Z = 0), then one edge per regulated bus pair
        self.parent, self.edge_in, self.children = {}, {}, {k: [] for k in range(c.nb)}
        self.edges = []                     # dicts: kind, Z, pm, and X slots of its current unknowns
        for l in range(c.nl):
            self.edges.append(dict(kind='line', tx=int(c.line_tx[l]), rx=int(c.line_rx[l]), Z=c.FZpu[l],
                                   pm=c.line_pm[l], cur=('line', l)))
        cnt = 0
        for t in range(c.ntf):
            slots = {}
            for ph in range(3):
                if c.xfm_pm[t][ph]:
                    slots[ph] = cnt
                    cnt += 1
            self.edges.append(dict(kind='xfm', tx=int(c.xfm_tx[t]), rx=int(c.xfm_rx[t]), Z=np.zeros((3, 3), complex),
                                   pm=c.xfm_pm[t], cur=('xfm', slots)))
        self.tf_lines = cnt
        vr_edges, cnt = {}, 0
        for m in range(c.nvr):                  # VR-phase counter in (VR, phase) order (:471-514)
            key = (int(c.vr_tx[m]), int(c.vr_reg[m]))
            e = vr_edges.setdefault(key, dict(kind='vr', tx=key[0], rx=key[1], pm=np.zeros(3, dtype=int),
                                              gamma=np.zeros(3), cur=('vr', {})))
            for ph in range(3):
                if c.vr_pm[m][ph]:
                    e['pm'][ph] = 1
                    e['gamma'][ph] = c.vr_gamma[m]
                    e['cur'][1][ph] = cnt       # in-leg slot 2*cnt, out-leg slot 2*cnt + 1
                    cnt += 1
        self.vr_lines = cnt
        self.edges += list(vr_edges.values())
        for ei, e in enumerate(self.edges):
            if e['rx'] in self.parent:
                raise ValueError(f"Network not radial. Bus {c.bus_names[e['rx']]} has two parents")
            self.parent[e['rx']], self.edge_in[e['rx']] = e['tx'], ei
            self.children[e['tx']].append(e['rx'])
        for k in range(1, c.nb):
            e = self.edges[self.edge_in[k]]
            if not np.array_equal(e['pm'], c.bus_pm[k]):
                raise NotImplementedError(f'bus {c.bus_names[k]}: phases differ from its incoming edge')
        self.N = 6 * (c.nb + c.nl) + 2 * self.tf_lines + 4 * self.vr_lines
        # per-phase constants of the second-order ZIP term (solution_nr3.py:228-261)
        self.cA, self.cB = np.zeros(3), np.zeros(3)
        for ph in range(3):
            A0, B0 = (1, 0) if ph == 0 else ((-1 / 2, -1 * np.sqrt(3) / 2) if ph == 1 else (-1 / 2, np.sqrt(3) / 2))
            h00 = -((A0 ** 2) * (A0 ** 2 + B0 ** 2) ** (-3 / 2)) + (A0 ** 2 + B0 ** 2) ** (-1 / 2)
            h11 = -((B0 ** 2) * (A0 ** 2 + B0 ** 2) ** (-3 / 2)) + ((A0 ** 2 + B0 ** 2) ** (-1 / 2))
            self.cA[ph] = c.aZ_p + (0.5 * c.aI_p * h00)
            self.cB[ph] = c.aZ_p + (0.5 * c.aI_p * h11)
        absent = [(k, ph) for k in range(1, c.nb) for ph in range(3) if c.bus_pm[k][ph] == 0]
        self.f0_absent = max([max(abs(VSLACK[ph].real), abs(VSLACK[ph].imag)) for _, ph in absent] + [0.0])

    def solve(self, spu=None, maxiter=MAXITER, wpu=None):
        """``wpu[S, nb, 3]`` real: per-scenario ``Solution.wpu`` as ``change_KCL_matrices`` enters it, subtracted
        from the constant term of both KCL rows of a bus-phase (solution_nr3.py:588)."""
        c = self.c
        if spu is None:
            spu = c.spu_matrix()[None]
        spu = np.asarray(spu, dtype=complex)
        S, nb, ne = spu.shape[0], c.nb, len(self.edges)
        V = np.tile(VSLACK[None, None, :], (S, nb, 1)) * c.bus_pm[None]     # present phases only
        V[:, 0] = VSLACK
        # Iin: current entering the child bus of an edge; Iout: current leaving its parent bus
        # (the same unknown for lines and transformers, the two legs of a regulator)
        Iin = np.zeros((S, ne, 3), dtype=complex)
        Iout = np.zeros((S, ne, 3), dtype=complex)
        iters = np.zeros(S, dtype=np.int32)
        active = np.ones(S, dtype=bool)
        first = True
        while active.any():
            idx = np.nonzero(active)[0]
            Vn, Iin_n, Iout_n, fmax = self._newton_step(V[idx], Iin[idx], Iout[idx], spu[idx], first,
                                                        None if wpu is None else np.asarray(wpu)[idx])
            V[idx], Iin[idx], Iout[idx] = Vn, Iin_n, Iout_n
            iters[idx] += 1
            with np.errstate(invalid='ignore'):
                active[idx] = (fmax >= 1e-9) & (iters[idx] < maxiter)
            first = False
        return dict(V=V, I=Iin[:, :c.nl], iterations=iters, XNR=self.to_xnr(V, Iin, Iout))

    def to_xnr(self, V, Iin, Iout):
        """Packed state -> the reference's XNR ordering (solution_nr3.py:46-61)."""
        c = self.c
        S, nb, nl = V.shape[0], c.nb, c.nl
        X = np.zeros((S, self.N))
        for ph in range(3):
            X[:, 2 * ph * nb:2 * ph * nb + 2 * nb:2] = V[:, :, ph].real
            X[:, 2 * ph * nb + 1:2 * ph * nb + 2 * nb:2] = V[:, :, ph].imag
            X[:, 6 * nb + 2 * ph * nl:6 * nb + 2 * ph * nl + 2 * nl:2] = Iin[:, :nl, ph].real
            X[:, 6 * nb + 2 * ph * nl + 1:6 * nb + 2 * ph * nl + 2 * nl:2] = Iin[:, :nl, ph].imag
        base_t = 6 * (nb + nl)
        base_r = base_t + 2 * self.tf_lines
        for ei, e in enumerate(self.edges):
            if e['kind'] == 'xfm':
                for ph, slot in e['cur'][1].items():
                    X[:, base_t + 2 * slot] = Iin[:, ei, ph].real
                    X[:, base_t + 2 * slot + 1] = Iin[:, ei, ph].imag
            elif e['kind'] == 'vr':
                for ph, cnt in e['cur'][1].items():
                    X[:, base_r + 2 * (2 * cnt)] = Iin[:, ei, ph].real
                    X[:, base_r + 2 * (2 * cnt) + 1] = Iin[:, ei, ph].imag
                    X[:, base_r + 2 * (2 * cnt + 1)] = Iout[:, ei, ph].real
                    X[:, base_r + 2 * (2 * cnt + 1) + 1] = Iout[:, ei, ph].imag
        return X

    @staticmethod
    def _blk(a, b, c_, d):
        """[S] x 4 -> [S,2,2]"""
        return np.stack([np.stack([a, b], -1), np.stack([c_, d], -1)], -2)

    def _newton_step(self, V, Iin, Iout, spu, first, wpu=None):
        c = self.c
        S = V.shape[0]
        beta_S = c.aPQ_p
        fmax = np.full(S, self.f0_absent if first else 0.0)
        Y = np.zeros((S, c.nb, 6, 6))
        a = np.zeros((S, c.nb, 6))
        aux = {}                              # per bus: what the forward sweep needs
        il = lambda z: np.stack([z.real, z.imag], axis=-1).reshape(z.shape[:-1] + (6,))   # [.,3] -> [.,6]
        upd = lambda f: np.maximum(fmax, np.abs(f).max(axis=1))
        for k in reversed(self.order[1:]):
            ei, p = self.edge_in[k], self.parent[k]
            e = self.edges[ei]
            pm = c.bus_pm[k].astype(float)
            v, i = V[:, k], Iin[:, ei]
            inet = i.copy()
            for ch in self.children[k]:
                inet = inet - Iout[:, self.edge_in[ch]]
            A, B = v.real, v.imag
            Cn, Dn = inet.real, inet.imag
            p_, q_ = spu[:, k].real, spu[:, k].imag
            cap = c.cappu[k]
            hAr, hBr, b0r = -p_ * self.cA, -p_ * self.cB, -p_ * beta_S
            hAi = -q_ * self.cA + cap * self.cA
            hBi = -q_ * self.cB + cap * self.cB
            b0i = (-q_ * beta_S) + (cap * beta_S)
            if wpu is not None:                                   # b = b_temp - wpu, both rows (:588)
                b0r, b0i = b0r - wpu[:, k], b0i - wpu[:, k]
            fr = (A * Cn + B * Dn + hAr * A ** 2 + hBr * B ** 2 + b0r) * pm      # KCL residual (:265-299)
            fi = (B * Cn - A * Dn + hAi * A ** 2 + hBi * B ** 2 + b0i) * pm
            r2 = np.stack([fr, fi], axis=-1).reshape(S, 6)
            fmax = upd(r2)
            # Dt = M^-1 D (2x2 per phase) + sum Y_c ; rt = M^-1 r2 + sum a_c
            Dt = np.zeros((S, 6, 6))
            rt = np.zeros((S, 6))
            for ph in range(3):
                if not pm[ph]:
                    continue
                a_, b_ = A[:, ph], B[:, ph]
                den = a_ * a_ + b_ * b_
                d00, d01 = Cn[:, ph] + 2 * hAr[:, ph] * a_, Dn[:, ph] + 2 * hBr[:, ph] * b_
                d10, d11 = -Dn[:, ph] + 2 * hAi[:, ph] * a_, Cn[:, ph] + 2 * hBi[:, ph] * b_
                # M^-1 = M / den, M = [[a, b], [b, -a]]
                Dt[:, 2 * ph, 2 * ph] = (a_ * d00 + b_ * d10) / den
                Dt[:, 2 * ph, 2 * ph + 1] = (a_ * d01 + b_ * d11) / den
                Dt[:, 2 * ph + 1, 2 * ph] = (b_ * d00 - a_ * d10) / den
                Dt[:, 2 * ph + 1, 2 * ph + 1] = (b_ * d01 - a_ * d11) / den
                rt[:, 2 * ph] = (a_ * fr[:, ph] + b_ * fi[:, ph]) / den
                rt[:, 2 * ph + 1] = (b_ * fr[:, ph] - a_ * fi[:, ph]) / den
            for ch in self.children[k]:
                Dt += Y[:, ch]
                rt += a[:, ch]
            if e['kind'] != 'vr':
                # KVL residual (:141-158; transformers :165-175 with Z = 0)
                Zr = _real_rep(e['Z'])
                r1 = il((V[:, p] * e['pm'] - v - np.einsum('ij,sj->si', e['Z'], i)) * pm)
                fmax = upd(r1)
                T = np.eye(6)[None] - Dt @ Zr
                E = np.repeat(e['pm'].astype(float), 2)
                rhs = np.concatenate([Dt * E[None, None, :], (rt + np.einsum('sij,sj->si', Dt, r1))[:, :, None]], axis=2)
                sol = np.linalg.solve(T, rhs)
                Y[:, k], a[:, k] = sol[:, :, :6], sol[:, :, 6]
                aux[k] = ('line', r1, Zr, E)
                continue
            # ---- regulator edge (:444-515): v_r = gamma v_m ; V_m conj(I_out) = V_r conj(I_in)
            vm, io = V[:, p], Iout[:, ei]
            ra = np.zeros((S, 6))            # G_reg rows: -gamma v_m + v_r
            fp = np.zeros((S, 6))            # H_reg rows (the reference's value carries a factor 2)
            Yk, ak = np.zeros((S, 6, 6)), np.zeros((S, 6))
            for ph in range(3):
                if not e['pm'][ph]:
                    continue
                g = e['gamma'][ph]
                sl = slice(2 * ph, 2 * ph + 2)
                A1, B1, A2, B2 = vm[:, ph].real, vm[:, ph].imag, A[:, ph], B[:, ph]
                Co, Do, Ci, Di = io[:, ph].real, io[:, ph].imag, i[:, ph].real, i[:, ph].imag
                ra[:, 2 * ph], ra[:, 2 * ph + 1] = -g * A1 + A2, -g * B1 + B2
                fp[:, 2 * ph] = 2 * ((A1 * Co + B1 * Do) - (A2 * Ci + B2 * Di))
                fp[:, 2 * ph + 1] = 2 * ((B1 * Co - A1 * Do) - (B2 * Ci - A2 * Di))
            fmax = upd(ra)
            fmax = upd(fp)
            u = rt - np.einsum('sij,sj->si', Dt, ra)          # di_in = u - Dt (gamma dv_m)
            G = np.repeat(e['gamma'], 2)                       # per-row gamma
            for ph in range(3):
                if not e['pm'][ph]:
                    continue
                sl = slice(2 * ph, 2 * ph + 2)
                A1, B1, A2, B2 = vm[:, ph].real, vm[:, ph].imag, A[:, ph], B[:, ph]
                Co, Do, Ci, Di = io[:, ph].real, io[:, ph].imag, i[:, ph].real, i[:, ph].imag
                Mm_inv = self._blk(A1, B1, B1, -A1) / (A1 * A1 + B1 * B1)[:, None, None]
                Mr = self._blk(A2, B2, B2, -A2)
                Nin, Nout = self._blk(Ci, Di, -Di, Ci), self._blk(Co, Do, -Do, Co)
                # a_k = Mm^-1 [F/2 + N_in r_a + M_r u] ;  Y_k = Mm^-1 [N_out - g N_in + M_r Dt G] (rows of this phase)
                rhs = fp[:, sl] / 2 + np.einsum('sij,sj->si', Nin, ra[:, sl]) + np.einsum('sij,sj->si', Mr, u[:, sl])
                ak[:, sl] = np.einsum('sij,sj->si', Mm_inv, rhs)
                row = np.einsum('sij,sjk->sik', Mr, Dt[:, sl, :] * G[None, None, :])
                row[:, :, sl] += Nout - e['gamma'][ph] * Nin
                Yk[:, sl, :] = np.einsum('sij,sjk->sik', Mm_inv, row)
            Y[:, k], a[:, k] = Yk, ak
            aux[k] = ('vr', ra, u, Dt, G)
        # slack rows (:84-109): dv_0 = v_0 - VSLACK
        dV = np.zeros((S, c.nb, 6))
        dV[:, 0] = il(V[:, 0] - VSLACK)
        fmax = upd(dV[:, 0])
        Vn, Iin_n, Iout_n = V.copy(), Iin.copy(), Iout.copy()
        cx = lambda r: r[:, 0::2] + 1j * r[:, 1::2]
        Vn[:, 0] = V[:, 0] - cx(dV[:, 0])
        for k in self.order[1:]:
            ei, p = self.edge_in[k], self.parent[k]
            dio = a[:, k] - np.einsum('sij,sj->si', Y[:, k], dV[:, p])     # current leaving the parent
            if aux[k][0] == 'line':
                _, r1, Zr, E = aux[k]
                dv = dV[:, p] * E - np.einsum('ij,sj->si', Zr, dio) - r1
                dii = dio
            else:
                _, ra, u, Dt, G = aux[k]
                dv = dV[:, p] * G + ra
                dii = u - np.einsum('sij,sj->si', Dt, dV[:, p] * G)
            dV[:, k] = dv
            Vn[:, k] = V[:, k] - cx(dv)
            Iin_n[:, ei] = Iin[:, ei] - cx(dii)
            Iout_n[:, ei] = Iout[:, ei] - cx(dio)
        return Vn, Iin_n, Iout_n, fmax

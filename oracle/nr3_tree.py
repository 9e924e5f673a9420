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
        if c.nvr or c.ntf:
            raise NotImplementedError('tree oracle: lines-only feeders')
        adj = _fbs_adjacency(c)
        self.order = _topo_sort(c.nb, adj)
        if self.order[0] != 0:
            raise NotImplementedError('slack bus (index 0) must be the root')
        self.parent, self.line_in, self.children = {}, {}, {k: [] for k in range(c.nb)}
        for l in range(c.nl):
            tx, rx = int(c.line_tx[l]), int(c.line_rx[l])
            self.parent[rx], self.line_in[rx] = tx, l
            self.children[tx].append(rx)
        for k in range(1, c.nb):
            l = self.line_in[k]
            if not np.array_equal(c.line_pm[l], c.bus_pm[k]):
                raise NotImplementedError(f'bus {c.bus_names[k]}: phases differ from its incoming line')
        # per-phase constants of the second-order ZIP term (solution_nr3.py:228-261)
        self.cA, self.cB = np.zeros(3), np.zeros(3)
        for ph in range(3):
            A0, B0 = (1, 0) if ph == 0 else ((-1 / 2, -1 * np.sqrt(3) / 2) if ph == 1 else (-1 / 2, np.sqrt(3) / 2))
            h00 = -((A0 ** 2) * (A0 ** 2 + B0 ** 2) ** (-3 / 2)) + (A0 ** 2 + B0 ** 2) ** (-1 / 2)
            h11 = -((B0 ** 2) * (A0 ** 2 + B0 ** 2) ** (-3 / 2)) + ((A0 ** 2 + B0 ** 2) ** (-1 / 2))
            self.cA[ph] = c.aZ_p + (0.5 * c.aI_p * h00)
            self.cB[ph] = c.aZ_p + (0.5 * c.aI_p * h11)
        self.Zr = np.array([_real_rep(c.FZpu[l]) for l in range(c.nl)])
        absent = [(k, ph) for k in range(1, c.nb) for ph in range(3) if c.bus_pm[k][ph] == 0]
        self.f0_absent = max([max(abs(VSLACK[ph].real), abs(VSLACK[ph].imag)) for _, ph in absent] + [0.0])

    def solve(self, spu=None, maxiter=MAXITER):
        c = self.c
        if spu is None:
            spu = c.spu_matrix()[None]
        spu = np.asarray(spu, dtype=complex)
        S, nb, nl = spu.shape[0], c.nb, c.nl
        V = np.tile(VSLACK[None, None, :], (S, nb, 1)) * c.bus_pm[None]     # present phases only
        I = np.zeros((S, nl, 3), dtype=complex)
        V[:, 0] = VSLACK
        iters = np.zeros(S, dtype=np.int32)
        active = np.ones(S, dtype=bool)
        first = True
        while active.any():
            idx = np.nonzero(active)[0]
            Vn, In, fmax = self._newton_step(V[idx], I[idx], spu[idx], first)
            V[idx], I[idx] = Vn, In
            iters[idx] += 1
            active[idx] = (fmax >= 1e-9) & (iters[idx] < maxiter)
            first = False
        return dict(V=V, I=I, iterations=iters, XNR=self.to_xnr(V, I))

    def to_xnr(self, V, I):
        c = self.c
        S, nb, nl = V.shape[0], c.nb, c.nl
        X = np.zeros((S, 6 * (nb + nl)))
        for ph in range(3):
            X[:, 2 * ph * nb:2 * ph * nb + 2 * nb:2] = V[:, :, ph].real
            X[:, 2 * ph * nb + 1:2 * ph * nb + 2 * nb:2] = V[:, :, ph].imag
            X[:, 6 * nb + 2 * ph * nl:6 * nb + 2 * ph * nl + 2 * nl:2] = I[:, :, ph].real
            X[:, 6 * nb + 2 * ph * nl + 1:6 * nb + 2 * ph * nl + 2 * nl:2] = I[:, :, ph].imag
        return X

    def _newton_step(self, V, I, spu, first):
        c = self.c
        S = V.shape[0]
        beta_S = c.aPQ_p
        fmax = np.full(S, self.f0_absent if first else 0.0)
        Y = np.zeros((S, c.nb, 6, 6))
        a = np.zeros((S, c.nb, 6))
        r1s = np.zeros((S, c.nb, 6))
        il = lambda z: np.stack([z.real, z.imag], axis=-1).reshape(z.shape[:-1] + (6,))   # [.,3] -> [.,6]
        for k in reversed(self.order[1:]):
            l, p = self.line_in[k], self.parent[k]
            pm = c.bus_pm[k].astype(float)
            v, i = V[:, k], I[:, l]
            inet = i.copy()
            for ch in self.children[k]:
                inet = inet - I[:, self.line_in[ch]]
            # KVL residual (:141-158)
            r1c = (V[:, p] * c.line_pm[l] - v - np.einsum('ij,sj->si', c.FZpu[l], i)) * pm
            r1 = il(r1c)
            A, B = v.real, v.imag
            Cn, Dn = inet.real, inet.imag
            p_, q_ = spu[:, k].real, spu[:, k].imag
            cap = c.cappu[k]
            hAr, hBr, b0r = -p_ * self.cA, -p_ * self.cB, -p_ * beta_S
            hAi = -q_ * self.cA + cap * self.cA
            hBi = -q_ * self.cB + cap * self.cB
            b0i = (-q_ * beta_S) + (cap * beta_S)
            fr = (A * Cn + B * Dn + hAr * A ** 2 + hBr * B ** 2 + b0r) * pm
            fi = (B * Cn - A * Dn + hAi * A ** 2 + hBi * B ** 2 + b0i) * pm
            r2 = np.stack([fr, fi], axis=-1).reshape(S, 6)
            fmax = np.maximum(fmax, np.maximum(np.abs(r1).max(axis=1), np.abs(r2).max(axis=1)))
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
            T = np.eye(6)[None] - Dt @ self.Zr[l]
            E = np.repeat(c.line_pm[l].astype(float), 2)
            rhs = np.concatenate([Dt * E[None, None, :], (rt + np.einsum('sij,sj->si', Dt, r1))[:, :, None]], axis=2)
            sol = np.linalg.solve(T, rhs)
            Y[:, k], a[:, k], r1s[:, k] = sol[:, :, :6], sol[:, :, 6], r1
        # slack rows (:84-109): dv_0 = v_0 - VSLACK
        dV = np.zeros((S, c.nb, 6))
        dV[:, 0] = il(V[:, 0] - VSLACK)
        fmax = np.maximum(fmax, np.abs(dV[:, 0]).max(axis=1))
        Vn, In = V.copy(), I.copy()
        cx = lambda r: r[:, 0::2] + 1j * r[:, 1::2]
        Vn[:, 0] = V[:, 0] - cx(dV[:, 0])
        for k in self.order[1:]:
            l, p = self.line_in[k], self.parent[k]
            E = np.repeat(c.line_pm[l].astype(float), 2)
            di = a[:, k] - np.einsum('sij,sj->si', Y[:, k], dV[:, p])
            dv = dV[:, p] * E - np.einsum('ij,sj->si', self.Zr[l], di) - r1s[:, k]
            dV[:, k] = dv
            Vn[:, k] = V[:, k] - cx(dv)
            In[:, l] = I[:, l] - cx(di)
        return Vn, In, fmax

"""Scenario studies on top of the batched solvers.

The reference runs a study one snapshot at a time: ``circuit.set_kW / set_kvar``
(``circuit.py:66-88``), ``solve()``, read the DataFrames (``solution.py:127-155``).
Here a study is a batch: its scenarios are synthesised on the device from small
tables, solved in chunks, and reduced on the device to what the study keeps per
scenario (losses, iteration counts, min / max |V|, voltage-band violations).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import pandas as pd

from . import _cabi


def voltage_summary(Vmag_rows, S: Optional[int] = None, vmin: float = 0.95, vmax: float = 1.05, stream=None):
    """Per-scenario (min |V|, max |V|, number of bus-phases outside [vmin, vmax]) of packed
    ``Vmag_rows`` ``[n_vrows, ld]`` (CUDA tensor; a NumPy array is uploaded) through
    ``gp_vmag_summary``.  Returns three CUDA tensors of length S."""
    torch = _cabi.require_cuda()
    if isinstance(Vmag_rows, np.ndarray):
        Vmag_rows = torch.from_numpy(np.ascontiguousarray(Vmag_rows, dtype=np.float64)).cuda()
    if Vmag_rows.dtype != torch.float64 or Vmag_rows.dim() != 2 or not Vmag_rows.is_contiguous():
        raise ValueError('Vmag_rows must be a contiguous [n_vrows, ld] float64 tensor')
    n_vrows, ld = int(Vmag_rows.shape[0]), int(Vmag_rows.shape[1])
    S = ld if S is None else int(S)
    dev = Vmag_rows.device
    lo = torch.empty(S, dtype=torch.float64, device=dev)
    hi = torch.empty(S, dtype=torch.float64, device=dev)
    cnt = torch.empty(S, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev) if stream is None else stream
    _cabi.check(_cabi.gp_vmag_summary(dev.index, S, ld, n_vrows, Vmag_rows.data_ptr(), float(vmin), float(vmax),
                                      lo.data_ptr(), hi.data_ptr(), cnt.data_ptr(), st.cuda_stream))
    return lo, hi, cnt


def scenario_frame(solution, result, param: str, scenario: int, orient: str = '') -> pd.DataFrame:
    """One scenario of a ``BatchResult`` as the DataFrame ``Solution.get_data_frame`` returns for a
    snapshot (``solution.py:127-155``): index = element names in OpenDSS order, columns A, B, C."""
    if param not in solution.SOLUTION_PARAMS:
        raise KeyError(param)
    element_group, cols, datatype = solution.SOLUTION_PARAMS[param]
    data = getattr(result, param)
    if data is None:
        raise ValueError(f'{param} was not among the outputs of this batch')
    index = getattr(solution.circuit, element_group).all_names()
    one = np.asarray(data[scenario])
    if type(solution).__name__ == 'SolutionNR3':       # NR3 keeps 3 x n (solution_nr3.py:634-663)
        one = one.T
    frame = pd.DataFrame(data=one[:len(index)], index=index, columns=cols, dtype=datatype)
    orient = orient or solution._orient
    return frame.transpose() if orient == 'cols' else frame


class DerStudy:
    """Hourly load profile x DER scenario study (BASELINE config C4).

    Scenario ``d * H + h`` is hour ``h`` of DER scenario ``d``:
    ``spu = m_h[h] * S_load - pen[d] * pv_h[h] * mask[d] * P_load`` per loaded bus-phase
    (SURVEY.md section 8(d), C4).  ``m_h``, ``pv_h``: ``[H]``; ``mask``: ``[D, n_srows]`` 0/1;
    ``pen``: ``[D]``.  Works with ``SolutionFBS`` and ``SolutionNR3``.
    """

    def __init__(self, solution, m_h, pv_h, mask, pen):
        self.sol = solution
        torch, dev = solution._dev()
        self.torch, self.dev = torch, dev
        t = solution.tables
        self.H, self.D = int(len(m_h)), int(len(pen))
        mask = np.ascontiguousarray(mask, dtype=np.float64)
        if mask.shape != (self.D, t.n_srows) or len(pv_h) != self.H:
            raise ValueError(f'mask must be [{self.D}, {t.n_srows}] and pv_h [{self.H}]')
        self.W0 = torch.from_numpy(np.ascontiguousarray(solution.load_weights().sum(axis=1))).to(dev)
        self.upload(m_h, pv_h, mask, pen)
        self.is_fbs = hasattr(solution, 'tolerance')
        self._host, self._spu = None, None

    def upload(self, m_h, pv_h, mask, pen):
        """(Re)load the hour and DER tables - the only inputs of the study that cross PCIe."""
        up = lambda a: self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)
        self.m_h, self.pv_h, self.mask, self.pen = up(m_h), up(pv_h), up(mask), up(pen)

    @property
    def n_scenarios(self) -> int:
        return self.H * self.D

    def spu_rows(self, first: int, S: int, out=None, stream=None, stride: int = 1):
        """Packed spu rows ``[n_srows, S]`` of scenarios ``first + s * stride``, s < S (CUDA tensor)."""
        torch, t = self.torch, self.sol.tables
        if out is None:
            out = torch.empty((t.n_srows, S), dtype=torch.complex128, device=self.dev)
        st = torch.cuda.current_stream(self.dev) if stream is None else stream
        _cabi.check(_cabi.gp_der_rows_strided(self.dev.index, S, int(out.shape[1]), int(first), int(stride), self.H,
                                              self.D, t.n_srows, self.W0.data_ptr(), self.m_h.data_ptr(),
                                              self.pv_h.data_ptr(), self.mask.data_ptr(), self.pen.data_ptr(),
                                              out.data_ptr(), st.cuda_stream))
        return out

    def run_chained(self, chain_hours: int = 24, first_chain: int = 0, chains: Optional[int] = None,
                    vmin: float = 0.95, vmax: float = 1.05) -> Dict[str, np.ndarray]:
        """The study as a time series: the ``H * D`` scenarios are cut into chains of
        ``chain_hours`` consecutive hours of one DER scenario (``H`` must be a multiple; a chain is a day by
        default); all chains advance together, one launch per hour of the chain, and hour k of a chain starts
        from the solution of hour k - 1 (``warm_start``) instead of the flat start - the first hour of
        every chain starts flat.  Same stop rule, so every result is a converged solution of the same
        equations within the solver tolerance, reached in fewer sweeps.  Returns the arrays of ``run``
        in scenario order (``chain * chain_hours + k``), for chains ``first_chain .. first_chain + chains``."""
        torch, sol, t = self.torch, self.sol, self.sol.tables
        L = int(chain_hours)
        if L < 1 or self.H % L:
            raise ValueError(f'chain_hours must divide H = {self.H}')
        n_all = self.n_scenarios // L
        chains = n_all - first_chain if chains is None else int(chains)
        if chains < 1 or first_chain < 0 or first_chain + chains > n_all:
            raise ValueError('chain range out of bounds')
        kinds = {'iterations': torch.int32, 'status': torch.int32, 'loss': torch.complex128,
                 'vmin': torch.float64, 'vmax': torch.float64, 'violations': torch.int32}
        dev_out = {k: torch.empty((chains, L), dtype=dt, device=self.dev) for k, dt in kinds.items()}
        spu = torch.empty((t.n_srows, chains), dtype=torch.complex128, device=self.dev)
        st = torch.cuda.current_stream(self.dev)
        state_outputs = ('V', 'I', 'Vmag', 'loss') if self.is_fbs else ('X', 'Vmag', 'loss')
        prev = None
        for k in range(L):
            self.spu_rows(first_chain * L + k, chains, out=spu, stream=st, stride=L)
            prev = sol.solve_batch(spu_rows=spu, outputs=state_outputs, return_device=True, stream=st,
                                   warm_start=prev)
            vlo, vhi, cnt = voltage_summary(prev.Vmag_rows, chains, vmin, vmax, stream=st)
            for key, src in (('iterations', prev.iterations), ('status', prev.status), ('loss', prev.loss),
                             ('vmin', vlo), ('vmax', vhi), ('violations', cnt)):
                dev_out[key][:, k].copy_(src[:chains])
        st.synchronize()
        return {k: v.reshape(-1).cpu().numpy() for k, v in dev_out.items()}

    def run(self, first: int = 0, count: Optional[int] = None, chunk: int = 131072, vmin: float = 0.95,
            vmax: float = 1.05) -> Dict[str, np.ndarray]:
        """Solve scenarios ``first .. first + count`` in chunks; per scenario keep iterations, status,
        loss, min |V|, max |V| and the number of bus-phases outside ``[vmin, vmax]`` (pinned host
        arrays; 36 bytes per scenario cross PCIe, nothing goes up but the tables)."""
        torch, sol, t = self.torch, self.sol, self.sol.tables
        count = self.n_scenarios - first if count is None else int(count)
        chunk = max(1, min(int(chunk), count))
        kinds = {'iterations': torch.int32, 'status': torch.int32, 'loss': torch.complex128,
                 'vmin': torch.float64, 'vmax': torch.float64, 'violations': torch.int32}
        if self._host is None or self._host['status'].numel() < count:      # pinning is slow: keep the buffers
            self._host = {k: torch.empty(count, dtype=dt).pin_memory() for k, dt in kinds.items()}
        host = {k: v[:count] for k, v in self._host.items()}
        if self._spu is None or self._spu.shape[1] != chunk:
            self._spu = torch.empty((t.n_srows, chunk), dtype=torch.complex128, device=self.dev)
        spu = self._spu
        if self.is_fbs:
            sol._maybe_specialise(chunk)
        st = torch.cuda.current_stream(self.dev)
        for lo in range(0, count, chunk):
            n = min(chunk, count - lo)
            self.spu_rows(first + lo, n, out=spu, stream=st)
            r = sol.solve_batch(spu_rows=spu[:, :n] if n == chunk else spu[:, :n].contiguous(),
                                outputs=('Vmag', 'loss'), return_device=True, stream=st)
            vlo, vhi, cnt = voltage_summary(r.Vmag_rows, n, vmin, vmax, stream=st)
            for key, src in (('iterations', r.iterations), ('status', r.status), ('loss', r.loss), ('vmin', vlo),
                             ('vmax', vhi), ('violations', cnt)):
                host[key][lo:lo + n].copy_(src[:n], non_blocking=True)
        st.synchronize()
        return {k: v.numpy() for k, v in host.items()}      # views of pinned buffers the next run() reuses


def volt_var_q(Vmag, points, minQ: float, maxQ: float):
    """``VoltVARController.get_Q`` (volt_var_controller.py:20-38) on a CUDA tensor of |V| (any shape),
    same arithmetic in the same order: maxQ below V1, a ramp to 0 at V2, 0 up to V3, a ramp to minQ at
    V4, minQ above; the sign of the result is flipped as the reference does."""
    import torch
    V1, V2, V3, V4 = (float(x) for x in points)
    lo_slope = (- maxQ) / (V2 - V1)
    lo_b = - (lo_slope * V2)
    hi_slope = (minQ) / (V4 - V3)
    hi_b = - (hi_slope * V3)
    q = torch.full_like(Vmag, float(minQ))
    q = torch.where(Vmag <= V4, hi_slope * Vmag + hi_b, q)
    q = torch.where(Vmag <= V3, torch.zeros_like(Vmag), q)
    q = torch.where(Vmag <= V2, lo_slope * Vmag + lo_b, q)
    q = torch.where(Vmag <= V1, torch.full_like(Vmag, float(maxQ)), q)
    return -q


class VoltVarStudy:
    """Quasi-static volt-var droop around the batched FBS or NR3 solve.

    One controller per listed (bus, phase): after every solve its reactive injection becomes
    ``wpu = get_Q(|V|)`` (``VoltVARController``, volt_var_controller.py:20-38; Q in pu) and the next
    solve starts from the last state (warm start) with that ``Solution.wpu`` - the loop a caller of the
    reference writes around ``solve()``, for every scenario of a batch at once.  Everything stays on the
    device between rounds; per round one solve launch, one gather of the controlled |V| rows and a
    handful of elementwise kernels."""

    def __init__(self, solution, controllers, points=(0.95, 0.98, 1.02, 1.05), minQ=-0.05, maxQ=0.05):
        self.sol = solution
        t = solution.tables
        names = solution.circuit.buses.all_names()
        self.ctl = []
        for bus, ph in controllers:
            b = names.index(bus) if isinstance(bus, str) else int(bus)
            row = int(t.vrow[b, ph])
            if row < 0:
                raise ValueError(f'bus {names[b]} has no phase {ph}')
            self.ctl.append(row)
        self.points, self.minQ, self.maxQ = tuple(points), float(minQ), float(maxQ)

    def run_in_iteration(self, spu_rows=None, load_mult=None, outputs=('V', 'Vmag'), **kw):
        """The droop INSIDE the FBS iteration: one solve in which every controlled
        bus-phase injects ``get_Q(|V|)`` of its current voltage whenever its power is formed
        (``GpSolveExtras.vv_*``).  Returns (BatchResult, Q ``[n_controllers, S]`` CUDA tensor: the injections at
        the returned voltages)."""
        torch, dev = self.sol._dev()
        t = self.sol.tables
        if hasattr(t, 'xnr_index'):
            raise NotImplementedError('the in-iteration droop is an FBS input; NR3 takes wpu rows (run())')
        names = self.sol.circuit.buses.all_names()
        ctl = []
        for row in self.ctl:
            b, p = np.argwhere(t.vrow == row)[0]
            ctl.append((int(b), int(p)))
        r = self.sol.solve_batch(spu_rows=spu_rows, load_mult=load_mult, outputs=tuple(set(outputs) | {'Vmag'}),
                                 return_device=True, volt_var=(ctl, self.points, self.minQ, self.maxQ), **kw)
        rows = torch.as_tensor(self.ctl, device=dev, dtype=torch.long)
        return r, volt_var_q(r.Vmag_rows.index_select(0, rows), self.points, self.minQ, self.maxQ)

    def run(self, spu_rows=None, load_mult=None, rounds: int = 4, outputs=('V', 'Vmag')):
        """``rounds`` solves; returns (last BatchResult, Q ``[rounds, n_controllers, S]`` CUDA tensor of the
        injections computed after each round, iterations ``[rounds, S]`` CUDA tensor)."""
        torch, dev = self.sol._dev()
        t = self.sol.tables
        rows = torch.as_tensor(self.ctl, device=dev, dtype=torch.long)
        # the state a warm start continues from: V and I rows (FBS), X rows (NR3, whose wpu enters the constant
        # term of the KCL rows, solution_nr3.py:588)
        state = {'X'} if hasattr(t, 'xnr_index') else {'V', 'I'}
        want = tuple(set(outputs) | state | {'Vmag'})
        prev, wpu, Q, its = None, None, [], []
        for _ in range(int(rounds)):
            prev = self.sol.solve_batch(spu_rows=spu_rows, load_mult=load_mult, outputs=want, return_device=True,
                                        warm_start=prev, wpu=wpu)
            q = volt_var_q(prev.Vmag_rows.index_select(0, rows), self.points, self.minQ, self.maxQ)
            if wpu is None:
                wpu = torch.zeros((t.n_vrows, prev.Vmag_rows.shape[1]), dtype=torch.float64, device=dev)
            wpu.index_copy_(0, rows, q)
            Q.append(q)
            its.append(prev.iterations.clone())
        return prev, torch.stack(Q), torch.stack(its)

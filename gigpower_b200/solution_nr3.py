"""Drop-in ``SolutionNR3`` backed by the sm_100a batched Newton-Raphson.

Same constructor, ``solve()``, ``map_XNR()`` and accessors as the reference
(``src/solution_nr3.py:28-35, 592-663``); results are stored ``3 x n`` like the
reference does.  ``solve_batch`` is the additive, data-parallel entry point.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _cabi
from .flatten import Nr3Tables, flatten_nr3, load_weights
from .solution import Solution
from .solution_fbs import _Feeder


@dataclass
class Nr3BatchResult:
    """Packed results ``[rows, S]``; ``unpack`` helpers give ``[S, 3, n]`` like the
    reference's per-snapshot ``3 x n`` matrices."""
    tables: Nr3Tables
    S: int
    iterations: np.ndarray
    status: np.ndarray
    fmax: Optional[np.ndarray] = None
    X_rows: Optional[object] = None
    V_rows: Optional[object] = None
    I_rows: Optional[object] = None
    Vmag_rows: Optional[object] = None
    sV_rows: Optional[object] = None
    iNode_rows: Optional[object] = None
    Stx_rows: Optional[object] = None
    Srx_rows: Optional[object] = None
    loss: Optional[object] = None

    @staticmethod
    def _host(x):
        return x.cpu().numpy() if hasattr(x, 'cpu') else np.asarray(x)

    @property
    def converged(self):
        return np.asarray(self._host(self.status)) == 0

    def _unpack(self, rows, rowmap, n):
        rows = self._host(rows)[:, :self.S]
        out = np.zeros((self.S, 3, n), dtype=rows.dtype)
        b, p = np.nonzero(rowmap[:n] >= 0)
        out[:, p, b] = rows[rowmap[b, p]].T
        return out

    @property
    def XNR(self):
        """``[S, N]`` in the reference's unknown order (solution_nr3.py:46-61)."""
        t = self.tables
        X = self._host(self.X_rows)[:, :self.S]
        out = np.zeros((self.S, t.N))
        out[:, t.xnr_index] = X.real.T
        out[:, t.xnr_index + 1] = X.imag.T
        return out

    V = property(lambda self: self._unpack(self.V_rows, self.tables.vrow, self.tables.vrow.shape[0]))
    Vmag = property(lambda self: self._unpack(self.Vmag_rows, self.tables.vrow, self.tables.vrow.shape[0]))
    sV = property(lambda self: self._unpack(self.sV_rows, self.tables.vrow, self.tables.vrow.shape[0]))
    i_Node = property(lambda self: self._unpack(self.iNode_rows, self.tables.vrow, self.tables.vrow.shape[0]))
    I = property(lambda self: self._unpack(self.I_rows, self.tables.irow, self.tables.n_lines))
    Stx = property(lambda self: self._unpack(self.Stx_rows, self.tables.irow, self.tables.n_lines))
    Srx = property(lambda self: self._unpack(self.Srx_rows, self.tables.irow, self.tables.n_lines))

    @property
    def losses(self):
        return self._host(self.loss)[:self.S]


class SolutionNR3(Solution):
    CONVERGENCE_TOLERANCE = 10 ** -6      # unused, as in the reference (solution_nr3.py:15, 611)

    @classmethod
    def set_zip_values(cls, zip_v):
        Solution.set_zip_values(zip_v)

    def __init__(self, dss_fp, **kwargs):
        super().__init__(dss_fp, **kwargs)
        self._set_orient('cols')
        self.tables: Nr3Tables = flatten_nr3(self.circuit)
        self._init_XNR()
        self._feeder = None
        self.itercount = 0

    def _init_XNR(self):                   # solution_nr3.py:37-82
        nb = self.circuit.buses.num_elements
        X = np.zeros((self.tables.N, 1), dtype=float)
        for ph in range(3):
            X[2 * ph * nb:2 * ph * nb + 2 * nb:2] = self.VSLACK[ph].real
            X[2 * ph * nb + 1:2 * ph * nb + 2 * nb:2] = self.VSLACK[ph].imag
        self.XNR = X

    def _dev(self):
        torch = _cabi.require_cuda()
        if self.device is None:
            self.device = torch.cuda.current_device()
        if self._feeder is None:
            desc, keep = _cabi.make_nr3_desc(self.tables)
            self._feeder = _Feeder(desc, keep, _cabi.gp_nr3_create, self.device)
        return torch, torch.device('cuda', self.device)

    def pack_spu(self, spu) -> np.ndarray:
        """``[S, nb, 3]`` (or ``[S, 3, nb]`` as this class orients its matrices is NOT
        accepted: give bus-major) complex -> packed ``[n_srows, S]``."""
        spu = np.asarray(spu, dtype=complex)
        if spu.ndim == 2:
            spu = spu[None]
        b, p = np.nonzero(self.tables.srow >= 0)
        rows = np.empty((self.tables.n_srows, spu.shape[0]), dtype=complex)
        rows[self.tables.srow[b, p]] = spu[:, b, p].T
        return rows

    def load_weights(self) -> np.ndarray:
        return load_weights(self.circuit, self.tables.srow, self.tables.n_srows)

    def solve_batch(self, spu=None, load_mult=None, spu_rows=None,
                    outputs=('X', 'V', 'Vmag', 'loss'), maxiter=None, return_device=False,
                    stream=None, taps=None, warm_start=None, wpu=None) -> Nr3BatchResult:
        """Newton-solve S scenarios from flat start.  Inputs as in
        ``SolutionFBS.solve_batch``; ``outputs`` from X, V, I, Vmag, sV, i_Node,
        Stx, Srx, loss.

        ``taps``: per-scenario regulator tap numbers (``gamma_rows_from_taps``).  ``warm_start``: an
        ``Nr3BatchResult`` of an earlier ``solve_batch(..., return_device=True)`` with X among its
        outputs and the same S: Newton starts from that X (as ``SolutionNR3.solve()`` called a second
        time starts from ``self.XNR``, solution_nr3.py:598) and updates the same tensor in place.
        ``wpu``: per-scenario ``Solution.wpu`` (``[S, nb, 3]`` real pu, or packed ``[n_vrows, S]`` float64 rows,
        NumPy or CUDA tensor): what a reference object computes after ``change_KCL_matrices`` has rewritten the
        constant term of its KCL rows with a non-zero ``circuit.get_wpu_matrix()`` (``b = ... - wpu`` on the real
        and the imaginary row of a bus-phase alike, solution_nr3.py:559-588); ``sV`` and ``i_Node`` carry
        ``+ 1j * wpu`` (nr3_lib/map_output.py:60)."""
        torch, dev = self._dev()
        t = self.tables
        maxiter = self.maxiter if maxiter is None else int(maxiter)
        if spu_rows is None:
            if load_mult is not None:
                spu_rows = self.load_weights() @ np.asarray(load_mult, dtype=float).T
            elif spu is not None:
                spu_rows = self.pack_spu(spu)
            else:
                orient = self.circuit._orient
                self.circuit._orient = 'rows'
                spu_rows = self.pack_spu(self.circuit.get_spu_matrix())
                self.circuit._orient = orient
        if isinstance(spu_rows, np.ndarray):
            d_spu = torch.from_numpy(np.ascontiguousarray(spu_rows, dtype=np.complex128)).to(dev)
        else:
            d_spu = spu_rows.to(device=dev, dtype=torch.complex128).contiguous()
        if d_spu.dim() != 2 or d_spu.shape[0] != t.n_srows:
            raise ValueError(f"spu_rows must be [{t.n_srows}, S], got {tuple(d_spu.shape)}")
        S = ld = int(d_spu.shape[1])
        h = self._feeder.handle
        want = set(outputs)
        c128, f64 = torch.complex128, torch.float64
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream() if stream is None else stream
            with torch.cuda.stream(st):
                if warm_start is not None:
                    X = warm_start.X_rows
                    if X is None or not hasattr(X, 'data_ptr') or tuple(X.shape) != (t.n_vrows + t.n_irows, ld):
                        raise ValueError("warm_start needs the X device tensor of an earlier solve_batch("
                                         "outputs=('X', ...), return_device=True) with the same S")
                else:
                    X = torch.empty((t.n_vrows + t.n_irows, ld), dtype=c128, device=dev)
                gamma = self.gamma_rows_from_taps(taps, S, dev) if taps is not None else None
                d_wpu = None
                if wpu is not None:
                    if isinstance(wpu, np.ndarray) or not hasattr(wpu, 'data_ptr'):
                        w = np.asarray(wpu, dtype=np.float64)
                        if w.ndim == 3:                    # [S, nb, 3] like Solution.wpu per scenario
                            b_, p_ = np.nonzero(t.vrow >= 0)
                            rows_w = np.zeros((t.n_vrows, w.shape[0]))
                            rows_w[t.vrow[b_, p_]] = w[:, b_, p_].T
                            w = rows_w
                        d_wpu = torch.from_numpy(np.ascontiguousarray(w)).to(dev)
                    else:
                        d_wpu = wpu.to(device=dev, dtype=f64).contiguous()
                    if tuple(d_wpu.shape) != (t.n_vrows, ld):
                        raise ValueError(f'wpu must be [S, nb, 3] or packed [{t.n_vrows}, {ld}], got {tuple(d_wpu.shape)}')
                extras = _cabi.make_extras(gamma, warm_start is not None, None, d_wpu)
                iters = torch.empty(S, dtype=torch.int32, device=dev)
                status = torch.empty(S, dtype=torch.int32, device=dev)
                fmax = torch.empty(S, dtype=f64, device=dev)
                wbytes = int(_cabi.gp_nr3_workspace_bytes(h, ld))
                if wbytes < 0:
                    _cabi.check(wbytes)
                work = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
                _cabi.check(_cabi.gp_nr3_solve_batch_ex(
                    h, S, ld, d_spu.data_ptr(), C.byref(extras) if extras is not None else None, X.data_ptr(),
                    iters.data_ptr(), status.data_ptr(), fmax.data_ptr(), work.data_ptr(), wbytes, maxiter,
                    st.cuda_stream))
                mk = lambda rows, dt: torch.zeros((max(rows, 1), ld), dtype=dt, device=dev)
                V = mk(t.n_vrows, c128) if 'V' in want else None
                I = mk(t.n_irows, c128) if 'I' in want else None
                Stx = mk(t.n_irows, c128) if 'Stx' in want else None
                Srx = mk(t.n_irows, c128) if 'Srx' in want else None
                sV = mk(t.n_vrows, c128) if 'sV' in want else None
                iN = mk(t.n_vrows, c128) if 'i_Node' in want else None
                Vmag = mk(t.n_vrows, f64) if 'Vmag' in want else None
                loss = torch.empty(S, dtype=c128, device=dev) if 'loss' in want else None
                if want & {'V', 'I', 'Stx', 'Srx', 'sV', 'i_Node', 'Vmag', 'loss'}:
                    p = lambda x: x.data_ptr() if x is not None else None
                    _cabi.check(_cabi.gp_nr3_map_output_ex(
                        h, S, ld, d_spu.data_ptr(), C.byref(extras) if extras is not None else None, X.data_ptr(), p(V),
                        p(I), p(Stx), p(Srx), p(sV), p(iN), p(Vmag), p(loss), st.cuda_stream))
                del work
        res = Nr3BatchResult(tables=t, S=S, iterations=iters, status=status, fmax=fmax,
                             X_rows=X if ('X' in want or warm_start is not None) else None, V_rows=V, I_rows=I, Vmag_rows=Vmag, sV_rows=sV,
                             iNode_rows=iN, Stx_rows=Stx, Srx_rows=Srx, loss=loss)
        if not return_device:
            st.synchronize()
            for k in ('iterations', 'status', 'fmax', 'X_rows', 'V_rows', 'I_rows', 'Vmag_rows', 'sV_rows',
                      'iNode_rows', 'Stx_rows', 'Srx_rows', 'loss'):
                v = getattr(res, k)
                if v is not None:
                    setattr(res, k, v.cpu().numpy())
        return res

    # ---- single snapshot (solution_nr3.py:592-663) -------------------------------
    def solve(self):
        r = self.solve_batch(outputs=('X', 'V', 'I', 'Vmag', 'sV', 'i_Node', 'Stx', 'Srx'))
        self._last = r
        self.itercount = int(r.iterations[0])      # the reference's local `itercount`
        self.XNR = r.XNR[0].reshape(-1, 1)
        self.converged = True                      # unconditional in the reference (:631)
        self.map_XNR()

    def change_KCL_matrices(self, der=0, capacitance=0, t=-1):
        """Not provided (solution_nr3.py:517-590).  The reference rewrites its dense
        H and b tensors in place with ONE scalar `der` / `capacitance` for every bus; this engine never builds
        those tensors.  The same studies are explicit per-scenario inputs here: ``solve_batch(spu=...)`` or
        ``solve_batch(load_mult=...)`` for loads and DER injections, ``solve_batch(wpu=...)`` for the reactive
        injection the method subtracts from the constant term (``- wpu``, :588), ``circuit.set_kW`` / ``set_kvar``
        for a single snapshot."""
        raise NotImplementedError(
            'change_KCL_matrices rewrites the dense NR3 tensors of the reference, which this engine does not build; '
            'pass per-scenario loads / DER injections to solve_batch(spu=... | load_mult=...) instead')

    def map_XNR(self):
        r = self._last
        self.V, self.I = r.V[0], r.I[0]
        self.Vmag = r.Vmag[0]
        self.Stx, self.Srx = r.Stx[0], r.Srx[0]
        self.i_Node, self.sV = r.i_Node[0], r.sV[0]

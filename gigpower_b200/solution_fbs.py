"""Drop-in ``SolutionFBS`` backed by the sm_100a batched Forward-Backward Sweep.

Same constructor, ``solve()`` and accessors as the reference
(``src/solution_fbs.py:27-136``); a single snapshot is a batch of one scenario
on the GPU.  ``solve_batch`` is the additive, data-parallel entry point.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _cabi
from .flatten import FbsTables, flatten_fbs, load_weights
from .solution import Solution


@dataclass
class BatchResult:
    """Results of a batched solve in packed row layout (``[rows, S]``); the
    ``unpack_*`` helpers give the reference's ``[S, n, 3]`` shapes."""
    tables: object
    S: int
    iterations: np.ndarray
    status: np.ndarray
    convergence_diff: Optional[np.ndarray] = None
    V_rows: Optional[object] = None
    I_rows: Optional[object] = None
    Vmag_rows: Optional[object] = None
    loss: Optional[object] = None
    sV_rows: Optional[object] = None
    Stx_rows: Optional[object] = None
    Srx_rows: Optional[object] = None

    @property
    def converged(self):
        return np.asarray(self.status) == 0

    @staticmethod
    def _host(x):
        return x.cpu().numpy() if hasattr(x, 'cpu') else np.asarray(x)

    def _unpack(self, rows, rowmap, n):
        rows = self._host(rows)[:, :self.S]
        out = np.zeros((self.S, n, 3), dtype=rows.dtype)
        b, p = np.nonzero(rowmap[:n] >= 0)
        out[:, b, p] = rows[rowmap[b, p]].T
        return out

    @property
    def V(self):
        return self._unpack(self.V_rows, self.tables.vrow, self.tables.vrow.shape[0])

    @property
    def Vmag(self):
        return self._unpack(self.Vmag_rows, self.tables.vrow, self.tables.vrow.shape[0])

    @property
    def sV(self):
        return self._unpack(self.sV_rows, self.tables.vrow, self.tables.vrow.shape[0])

    @property
    def I(self):
        return self._unpack(self.I_rows, self.tables.irow, self.tables.irow.shape[0])

    @property
    def Stx(self):
        return self._unpack(self.Stx_rows, self.tables.irow, self.tables.n_lines)

    @property
    def Srx(self):
        return self._unpack(self.Srx_rows, self.tables.irow, self.tables.n_lines)

    @property
    def losses(self):
        return self._host(self.loss)[:self.S]


class _Feeder:
    """Owns the device handle of one flattened feeder."""

    def __init__(self, desc, keep, create, device):
        self.handle = C.c_void_p()
        self._keep = keep
        _cabi.check(create(C.byref(desc), int(device), C.byref(self.handle)))

    def __del__(self):
        try:
            if self.handle:
                _cabi.gp_feeder_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class SolutionFBS(Solution):

    @classmethod
    def set_zip_values(cls, zip_v):
        Solution.set_zip_values(zip_v)

    def __init__(self, dss_fp, **kwargs):
        super().__init__(dss_fp, **kwargs)
        self.tables: FbsTables = flatten_fbs(self.circuit)
        self.adj, self.reverse_adj = self.tables.adj, self.tables.reverse_adj
        self.topo_order = self.tables.topo_order
        self.root = self.circuit.buses.get_element(self.topo_order[0])
        self.Vref = np.array([1, np.exp(1j * 240 * np.pi / 180), np.exp(1j * 120 * np.pi / 180)],
                             dtype=complex)
        self.tolerance = abs((self.Vref[1]) * 10 ** -6)
        self._feeder = None
        self._specialised = None      # None = not tried, True / False = outcome of specialise()

    #: batches of at least this many scenarios compile the feeder-specialised kernel (about one
    #: second with NVRTC, once per instance) when the feeder fits the on-chip kernel
    SPECIALISE_MIN_BATCH = 8192
    #: ... and at least this many 32-scenario warps of its state fit in one SM's shared memory
    #: (measured on B200: 34-bus, 2 warps: 1.4x faster than streaming; 37-bus, 1 warp: slower)
    SPECIALISE_MIN_WARPS = 2

    # ---- device plumbing -----------------------------------------------------
    def _dev(self):
        torch = _cabi.require_cuda()
        if self.device is None:
            self.device = torch.cuda.current_device()
        if self._feeder is None:
            desc, keep = _cabi.make_fbs_desc(self.tables)
            self._feeder = _Feeder(desc, keep, _cabi.gp_fbs_create, self.device)
        return torch, torch.device('cuda', self.device)

    def specialise(self) -> bool:
        """Generate, compile (NVRTC) and load the feeder-specialised on-chip kernel
        (``gp_fbs_specialise``).  Returns False - and the table-driven kernels keep being used -
        when the feeder does not fit in shared memory or no runtime compiler is available."""
        self._dev()
        if self._specialised is None:
            rc = _cabi.gp_fbs_specialise(self._feeder.handle)
            self._specialised = rc == 0
            if rc != 0:
                import warnings
                try:
                    _cabi.check(rc)
                except _cabi.GpError as e:
                    warnings.warn(f'feeder-specialised FBS kernel unavailable, using the table-driven kernel: {e}')
        return self._specialised

    def _maybe_specialise(self, S):
        if (self._specialised is None and S >= self.SPECIALISE_MIN_BATCH
                and _cabi.gp_fbs_jit_warps(self._feeder.handle) >= self.SPECIALISE_MIN_WARPS):
            self.specialise()

    def pack_spu(self, spu) -> np.ndarray:
        """``[S, nb, 3]`` complex (the reference's ``Solution.spu`` per scenario)
        -> packed ``[n_srows, S]``."""
        spu = np.asarray(spu, dtype=complex)
        if spu.ndim == 2:
            spu = spu[None]
        b, p = np.nonzero(self.tables.srow >= 0)
        rows = np.empty((self.tables.n_srows, spu.shape[0]), dtype=complex)
        rows[self.tables.srow[b, p]] = spu[:, b, p].T
        return rows

    def load_weights(self) -> np.ndarray:
        """``[n_srows, n_loads]`` complex W with ``spu_rows = W @ mult.T`` for
        per-load multipliers (load.py:16-24 split per phase, summed per bus)."""
        return load_weights(self.circuit, self.tables.srow, self.tables.n_srows)

    # ---- batched solve -------------------------------------------------------
    def solve_batch(self, spu=None, load_mult=None, spu_rows=None, outputs=('V', 'Vmag', 'loss'),
                    maxiter=None, tol=None, return_device=False, stream=None, taps=None,
                    warm_start=None, gather=None, wpu=None, volt_var=None) -> BatchResult:
        """Solve S scenarios from flat start.

        Give ONE of: ``spu`` ``[S, nb, 3]`` complex; ``load_mult`` ``[S, n_loads]``
        multipliers on every load's kW and kvar; ``spu_rows`` packed
        ``[n_srows, S]`` complex128 (NumPy, or a CUDA torch tensor for inputs
        already resident in HBM).  ``outputs`` selects what is produced besides
        iterations/status: any of V, I, Vmag, loss, sV, Stx, Srx.

        ``taps``: per-scenario regulator tap numbers (see ``gamma_rows_from_taps``); scenario s is
        then what a reference object built with ``taps[s]`` computes.  ``warm_start``: a
        ``BatchResult`` of an earlier ``solve_batch(..., outputs=(... 'V', 'I'), return_device=True)``
        with the same number of scenarios; the sweep continues from its V and I (what the reference
        does when ``Vtest`` and ``iterations`` are reset and ``solve()`` is called again) instead of
        the flat start, and the result's V/I tensors are the SAME device tensors, updated in place.
        ``gather``: ``(PeerGather, slot)`` (``gigpower_b200.sharding``): the solve also stores this rank's
        |V| rows and losses into that slot of every rank's gather buffer (the all-gather of SURVEY.md
        8(e), fused into the kernel epilogue); read them with ``PeerGather.vmag_all / loss_all`` after
        ``PeerGather.wait()``.  ``(PeerGather, slot, 'own')`` stores into this rank's own buffer only,
        for ``PeerGather.push_own`` to distribute with the copy engines.  The batch must have the
        PeerGather's per-rank size.
        ``wpu``: per-scenario ``Solution.wpu`` - the controllable reactive injection ``+ 1j * wpu`` of
        ``calc_sV`` (solution.py:106, 203), zero in the reference unless the caller sets the attribute:
        ``[S, nb, 3]`` real pu, or packed ``[n_vrows, S]`` float64 rows (NumPy or CUDA tensor).
        ``volt_var``: ``(controllers, points, minQ, maxQ)`` - volt-var droop INSIDE the iteration: at every
        ``(bus name or index, phase)`` of ``controllers`` the injection is ``wpu = VoltVARController.get_Q(|V|)``
        (volt_var_controller.py:20-38, breakpoints ``points = (V1, V2, V3, V4)``, ``minQ``/``maxQ`` in pu),
        re-evaluated from the bus's current voltage whenever its power is formed - what the reference's
        ``SolutionFBS`` computes when its own controller objects are consulted before each ``calc_sV``.
        """
        torch, dev = self._dev()
        t = self.tables
        maxiter = self.maxiter if maxiter is None else int(maxiter)
        tol = self.tolerance if tol is None else float(tol)
        if spu_rows is None:
            if load_mult is not None:
                spu_rows = self.load_weights() @ np.asarray(load_mult, dtype=float).T
            elif spu is not None:
                spu_rows = self.pack_spu(spu)
            else:
                spu_rows = self.pack_spu(self.spu)
        if isinstance(spu_rows, np.ndarray):
            d_spu = torch.from_numpy(np.ascontiguousarray(spu_rows, dtype=np.complex128)).to(dev)
        else:
            d_spu = spu_rows.to(device=dev, dtype=torch.complex128).contiguous()
        if d_spu.dim() != 2 or d_spu.shape[0] != t.n_srows:
            raise ValueError(f"spu_rows must be [{t.n_srows}, S], got {tuple(d_spu.shape)}")
        S = ld = int(d_spu.shape[1])
        self._maybe_specialise(S)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream() if stream is None else stream
            with torch.cuda.stream(st):
                if warm_start is not None:
                    V, I = warm_start.V_rows, warm_start.I_rows
                    if (V is None or I is None or not hasattr(V, 'data_ptr') or tuple(V.shape) != (t.n_vrows, ld)
                            or tuple(I.shape) != (max(t.n_irows, 1), ld)):
                        raise ValueError('warm_start needs the V and I device tensors of an earlier solve_batch('
                                         "outputs=(..., 'V', 'I'), return_device=True) with the same S")
                else:
                    V = torch.empty((t.n_vrows, ld), dtype=torch.complex128, device=dev)
                    I = torch.empty((max(t.n_irows, 1), ld), dtype=torch.complex128, device=dev)
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
                        d_wpu = wpu.to(device=dev, dtype=torch.float64).contiguous()
                    if tuple(d_wpu.shape) != (t.n_vrows, ld):
                        raise ValueError(f'wpu must be [S, nb, 3] or packed [{t.n_vrows}, {ld}], got {tuple(d_wpu.shape)}')
                peers = None
                if gather is not None:
                    pg, slot = gather[0], gather[1]
                    if (pg.n_vrows, pg.ld) != (t.n_vrows, ld):
                        raise ValueError(f'gather buffers are for [{pg.n_vrows}, {pg.ld}] blocks, this batch is '
                                         f'[{t.n_vrows}, {ld}]')
                    own = len(gather) > 2 and gather[2] == 'own'
                    peers = pg.own_destination(slot) if own else pg.destinations(slot)
                vv = None
                if volt_var is not None:
                    ctl, points, minQ, maxQ = volt_var
                    names = self.circuit.buses.all_names()
                    flags = np.zeros(t.n_vrows, dtype=np.int32)
                    for bus, ph in ctl:
                        b_i = names.index(bus) if isinstance(bus, str) else int(bus)
                        if t.vrow[b_i, int(ph)] < 0:
                            raise ValueError(f'bus {names[b_i]} has no phase {ph}')
                        flags[t.vrow[b_i, int(ph)]] = 1
                    vv = (torch.from_numpy(flags).to(dev), tuple(points), float(minQ), float(maxQ))
                extras = _cabi.make_extras(gamma, warm_start is not None, peers, d_wpu, vv)
                iters = torch.empty(S, dtype=torch.int32, device=dev)
                status = torch.empty(S, dtype=torch.int32, device=dev)
                diff = torch.empty(S, dtype=torch.float64, device=dev)
                want = set(outputs)
                mk = lambda rows, dt: torch.empty((max(rows, 1), ld), dtype=dt, device=dev)
                sV = mk(t.n_vrows, torch.complex128) if 'sV' in want else None
                Vmag = mk(t.n_vrows, torch.float64) if 'Vmag' in want else None
                Stx = torch.zeros((max(t.n_irows, 1), ld), dtype=torch.complex128, device=dev) if 'Stx' in want else None
                Srx = torch.zeros((max(t.n_irows, 1), ld), dtype=torch.complex128, device=dev) if 'Srx' in want else None
                loss = torch.empty(S, dtype=torch.complex128, device=dev) if 'loss' in want else None
                p = lambda x: x.data_ptr() if x is not None else None
                # solve + |V| + loss in one call (one launch with the feeder-specialised kernel)
                _cabi.check(_cabi.gp_fbs_solve_batch_ex(
                    self._feeder.handle, S, ld, d_spu.data_ptr(), C.byref(extras) if extras is not None else None,
                    V.data_ptr(), I.data_ptr(), iters.data_ptr(), status.data_ptr(), diff.data_ptr(), p(Vmag),
                    p(loss), tol, maxiter, st.cuda_stream))
                if want & {'sV', 'Stx', 'Srx'}:
                    _cabi.check(_cabi.gp_fbs_post_ex(
                        self._feeder.handle, S, ld, d_spu.data_ptr(), C.byref(extras) if extras is not None else None,
                        V.data_ptr(), I.data_ptr(), p(sV), None, p(Stx), p(Srx), None, st.cuda_stream))
        res = BatchResult(tables=t, S=S, iterations=iters, status=status, convergence_diff=diff,
                          V_rows=V if ('V' in want or warm_start is not None) else None,
                          I_rows=I if ('I' in want or warm_start is not None) else None,
                          Vmag_rows=Vmag, loss=loss, sV_rows=sV, Stx_rows=Stx, Srx_rows=Srx)
        if not return_device:
            st.synchronize()
            for k in ('iterations', 'status', 'convergence_diff', 'V_rows', 'I_rows', 'Vmag_rows', 'loss',
                      'sV_rows', 'Stx_rows', 'Srx_rows'):
                v = getattr(res, k)
                if v is not None:
                    setattr(res, k, v.cpu().numpy())
        return res

    def solve_batch_host(self, spu_rows, V_out=None, I_out=None, Vmag_out=None, loss_out=None,
                         maxiter=None, tol=None):
        """Host-buffer path through ``gp_fbs_solve_batch_host``: ``spu_rows`` is a
        (preferably pinned) CPU torch tensor / NumPy array ``[n_srows, S]``
        complex128; outputs are written into the given host arrays.  Returns
        (iterations, status, convergence_diff) as NumPy views of pinned buffers that the next
        call reuses."""
        torch, dev = self._dev()
        t = self.tables
        maxiter = self.maxiter if maxiter is None else int(maxiter)
        tol = self.tolerance if tol is None else float(tol)

        def ptr(x):
            if x is None:
                return None
            return x.data_ptr() if hasattr(x, 'data_ptr') else x.ctypes.data

        S = ld = int(spu_rows.shape[1])
        self._maybe_specialise(S)
        # pinned result vectors: a D2H copy into pageable memory would block the host inside
        # the library and serialise the copy/compute pipeline
        if getattr(self, '_host_vecs', None) is None or self._host_vecs[0].numel() < S:
            self._host_vecs = (torch.empty(S, dtype=torch.int32).pin_memory(),
                               torch.empty(S, dtype=torch.int32).pin_memory(),
                               torch.empty(S, dtype=torch.float64).pin_memory())
        iters, status, diff = (v[:S].numpy() for v in self._host_vecs)
        _cabi.check(_cabi.gp_fbs_solve_batch_host(
            self._feeder.handle, S, ld, ptr(spu_rows), ptr(V_out), ptr(I_out), ptr(Vmag_out), ptr(loss_out),
            iters.ctypes.data, status.ctypes.data, diff.ctypes.data, tol, maxiter))
        return iters, status, diff

    # ---- derived results of the CURRENT snapshot state (solution.py:177-233) --------
    def _post_snapshot(self, want):
        """``self.V`` / ``self.I`` / ``self.spu`` / ``self.wpu`` as they stand (``n x 3``, the caller may have
        changed them) through ``gp_fbs_post_ex`` as a batch of one scenario: the same kernels that
        produce sV, |V|, Stx, Srx for a batch.  Returns the unpacked ``n x 3`` arrays asked for."""
        torch, dev = self._dev()
        t = self.tables

        def pack(mat, rowmap, n_rows, dtype):
            rows = np.zeros((max(n_rows, 1), 1), dtype=dtype)
            b, p = np.nonzero(rowmap >= 0)
            rows[rowmap[b, p], 0] = np.asarray(mat)[b, p]
            return torch.from_numpy(rows).to(dev)

        V = pack(self.V, t.vrow, t.n_vrows, complex)
        I = pack(self.I, t.irow, t.n_irows, complex)
        spu = torch.from_numpy(self.pack_spu(self.spu)).to(dev) if t.n_srows else torch.zeros((1, 1), dtype=torch.complex128, device=dev)
        wpu = pack(self.wpu, t.vrow, t.n_vrows, float) if np.any(self.wpu) else None
        ex = _cabi.make_extras(wpu=wpu)
        mk = lambda rows, dt: torch.zeros((max(rows, 1), 1), dtype=dt, device=dev)
        out = {'sV': mk(t.n_vrows, torch.complex128) if 'sV' in want else None,
               'Vmag': mk(t.n_vrows, torch.float64) if 'Vmag' in want else None,
               'Stx': mk(t.n_irows, torch.complex128) if 'Stx' in want else None,
               'Srx': mk(t.n_irows, torch.complex128) if 'Srx' in want else None}
        p = lambda x: x.data_ptr() if x is not None else None
        st = torch.cuda.current_stream(dev)
        _cabi.check(_cabi.gp_fbs_post_ex(self._feeder.handle, 1, 1, spu.data_ptr(),
                                         C.byref(ex) if ex is not None else None, V.data_ptr(), I.data_ptr(),
                                         p(out['sV']), p(out['Vmag']), p(out['Stx']), p(out['Srx']), None, st.cuda_stream))
        st.synchronize()
        res = BatchResult(tables=t, S=1, iterations=None, status=None, sV_rows=out['sV'], Vmag_rows=out['Vmag'],
                          Stx_rows=out['Stx'], Srx_rows=out['Srx'])
        return {k: getattr(res, k)[0] for k in want}

    def calc_sV(self, bus=None):
        """``sV = spu (aPQ + aI |V| + aZ |V|^2) - j cappu |V|^2 + j wpu`` from the current ``self.V``, zero on
        absent phases (solution.py:177-220); with ``bus`` only that bus's row of ``self.sV`` is updated."""
        sV = self._post_snapshot(('sV',))['sV']
        if bus:
            i = self.circuit.buses.get_idx(bus)
            self.sV[i] = sV[i]
        else:
            self.sV = sV

    def calc_Vmag(self):
        """``Vmag = abs(V)`` (solution.py:232-233)."""
        self.Vmag = self._post_snapshot(('Vmag',))['Vmag']

    def calc_Stx(self):
        """``Stx = V[tx] conj(I)`` of the real lines (solution.py:222-225)."""
        self.Stx = self._post_snapshot(('Stx',))['Stx']

    def calc_Srx(self):
        """``Srx = V[rx] conj(I)`` of the real lines (solution.py:227-230)."""
        self.Srx = self._post_snapshot(('Srx',))['Srx']

    # ---- single snapshot (solution_fbs.py:76-136) -------------------------------
    def solve(self):
        t = self.tables
        if not (self.iterations > 0 and (self.converged or self.iterations >= self.maxiter)):
            r = self.solve_batch(spu=self.spu, outputs=('V', 'I', 'Vmag', 'sV', 'Stx', 'Srx'))
            self._last = r
            self.iterations = int(r.iterations[0])
            self.convergence_diff = float(r.convergence_diff[0])
            self.converged = bool(r.status[0] == 0)
        r = self._last
        self.V = r.V[0]
        self.I = r.I[0]
        self.sV = r.sV[0]
        self.Vmag = r.Vmag[0]
        self.Stx = r.Stx[0]
        self.Srx = r.Srx[0]
        self.Vtest = self.V[self.circuit.buses.get_idx(self.root)]

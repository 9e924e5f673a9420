"""Scenario sharding across GPUs (one process per GPU, ``torch.distributed``).

Scenarios are independent given the replicated, kilobyte-scale feeder tables,
so the batch is cut into contiguous blocks, one per rank, and solved with no
traffic at all.  The only exchange is the final gather of per-scenario
voltage magnitudes and losses; it is an all-gather over
NCCL on GPUs and works unchanged over gloo on CPU tensors, which is how the
host-side logic is tested without a GPU.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np


def shard_range(S: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank``; sizes differ by at most one,
    the first ``S % world`` ranks get the extra scenario."""
    if world <= 0 or not (0 <= rank < world) or S < 0:
        raise ValueError(f'bad shard request S={S} rank={rank} world={world}')
    base, extra = divmod(S, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def bind_to_gpu_numa(device_index: int) -> Optional[list]:
    """Pin this process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA node / PCIe root),
    so that the pinned host buffers it allocates afterwards are first-touched on that node and its copies do not
    cross the socket interconnect.  One process per GPU makes this a per-rank setting.  Returns the core list, or
    None when NVML or the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = device_index
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            try:
                idx = int(vis.split(',')[device_index])
            except ValueError:
                pass
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cores = [64 * w + b for w, m in enumerate(masks) for b in range(64) if (int(m) >> b) & 1]
        cores = [c for c in cores if c in os.sched_getaffinity(0)]
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None


def gather_rows(local, S: int, group=None):
    """All-gather a per-scenario array whose LAST dimension is this rank's
    scenario block (``[..., n_local]``) into ``[..., S]`` on every rank.
    Blocks may be ragged: they are padded to the largest block for the
    collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(S, rank, world)
    if local.shape[-1] != hi - lo:
        raise ValueError(f'rank {rank}: expected a block of {hi - lo} scenarios, got {local.shape[-1]}')
    nmax = -(-S // world) if S else 0
    pad = torch.zeros(local.shape[:-1] + (nmax,), dtype=local.dtype, device=local.device)
    pad[..., :hi - lo] = local
    out = torch.empty((world,) + tuple(pad.shape), dtype=local.dtype, device=local.device)
    src = torch.view_as_real(pad.contiguous()) if pad.is_complex() else pad.contiguous()   # gloo: no complex
    dst = torch.view_as_real(out) if out.is_complex() else out
    dist.all_gather_into_tensor(dst.view(-1), src.view(-1), group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(S, r, world)
        parts.append(out[r][..., :b - a])
    return torch.cat(parts, dim=-1)


def solve_sharded(solution, spu_rows: np.ndarray, outputs=('Vmag', 'loss'), group=None) -> Dict[str, object]:
    """Solve the scenarios of ``spu_rows`` (``[n_srows, S]``, the same array on
    every rank, or at least this rank's block of it) on this rank's GPU and
    gather ``Vmag`` ``[n_vrows, S]``, ``loss`` ``[S]``, ``iterations`` ``[S]`` and
    ``status`` ``[S]`` on every rank.  Works for SolutionFBS and SolutionNR3."""
    import torch.distributed as dist
    S = int(spu_rows.shape[1])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(S, rank, world)
    res = solution.solve_batch(spu_rows=np.ascontiguousarray(spu_rows[:, lo:hi]),
                               outputs=tuple(outputs), return_device=True)
    out = {'iterations': gather_rows(res.iterations, S, group), 'status': gather_rows(res.status, S, group)}
    if 'Vmag' in outputs:
        out['Vmag'] = gather_rows(res.Vmag_rows, S, group)
    if 'loss' in outputs:
        out['loss'] = gather_rows(res.loss, S, group)
    return out


class _DevArray:
    """Raw device memory as something ``torch.as_tensor`` wraps without a copy."""

    def __init__(self, ptr, shape, typestr, strides=None):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 3, 'strides': tuple(strides) if strides else None}


class PeerGather:
    """The path's one exchange without a collective kernel.

    Every rank owns ``slots`` gather buffers, each made of ``world`` contiguous blocks - block r =
    rank r's ``[n_vrows][ld]`` float64 |V| rows followed by its ``[ld]`` complex128 losses -
    allocated through ``gp_peer_alloc`` and mapped into every
    other rank's process with CUDA IPC (``gp_peer_open``; the handles travel through
    ``torch.distributed.all_gather_object`` - NCCL or gloo, control plane only).  A rank's solve
    then writes block ``rank`` of the chosen slot on EVERY rank: either the solve kernel's own
    epilogue does it (``solve_batch(gather=(pg, slot))``, stores over NVLink / NVSwitch that
    overlap the next tile), or the copy engines move finished rows (``push``; ``push_own`` when the
    solve wrote this rank's block of its own buffer, ``gather=(pg, slot, 'own')``).  Nothing
    competes with the solve for SMs, which is what an NCCL all-gather next to a persistent
    one-block-per-SM kernel does (measured on 8 x B200: the two serialise).

    ``wait()`` makes a slot readable: this rank's stream work is complete and so is every
    other rank's (barrier).  Re-using a slot needs the same wait on the reader's side first.
    """

    def __init__(self, n_vrows: int, ld: int, slots: int = 1, group=None, device=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _cabi
        self._cabi, self._C, self._torch = _cabi, C, torch
        self.group = group
        multi = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        if self.world > _cabi.GP_MAX_PEERS:
            raise ValueError(f'at most {_cabi.GP_MAX_PEERS} ranks')
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.n_vrows, self.ld, self.slots = int(n_vrows), int(ld), int(slots)
        self.vmag_bytes = self.n_vrows * self.ld * 8              # per rank
        self.block_bytes = self.vmag_bytes + self.ld * 16         # one rank's |V| rows + losses
        self.slot_bytes = (self.world * self.block_bytes + 255) // 256 * 256
        ptr = C.c_void_p()
        handle = C.create_string_buffer(_cabi.GP_PEER_HANDLE_BYTES)
        self._local, err = None, None
        try:
            _cabi.check(_cabi.gp_peer_alloc(self.device, self.slot_bytes * self.slots, C.byref(ptr), handle))
            self._local = ptr.value
        except _cabi.GpError as e:
            if self.world == 1:
                raise
            err = str(e)
        self.bases = [None] * self.world                 # base address of every rank's buffer in THIS process
        self.bases[self.rank] = self._local
        if self.world > 1:
            # failures are made collective: every rank takes part in both exchanges, then all raise or none
            handles = [None] * self.world
            dist.all_gather_object(handles, (self.rank, bytes(handle.raw) if err is None else None, err), group=group)
            errs = [f'rank {r}: {e}' for r, h, e in handles if h is None]
            if not errs:
                try:
                    for r, h, _ in handles:
                        if r == self.rank:
                            continue
                        p = C.c_void_p()
                        _cabi.check(_cabi.gp_peer_open(self.device, C.create_string_buffer(h, len(h)), C.byref(p)))
                        self.bases[r] = p.value
                except _cabi.GpError as e:
                    err = str(e)
            opened = [None] * self.world
            dist.all_gather_object(opened, (self.rank, err), group=group)
            errs += [f'rank {r}: {e}' for r, e in opened if e is not None]
            if errs:
                for r, b in enumerate(self.bases):
                    if r != self.rank and b is not None:
                        _cabi.gp_peer_close(self.device, b)
                dist.barrier(group=group)
                if self._local is not None:
                    _cabi.gp_peer_free(self.device, self._local)
                self._local = None
                raise _cabi.GpError('peer-mapped gather buffers unavailable: ' + '; '.join(sorted(set(errs))))

    # ---- addresses -----------------------------------------------------------------------
    def destinations(self, slot: int):
        """(|V| addresses, loss addresses): block ``rank`` of ``slot`` in every rank's buffer, in the
        order rank + 1, rank + 2, ..., rank: at any moment the ranks of a box then write to DIFFERENT
        peers (a permutation) instead of all converging on rank 0's NVLink ingress first."""
        if not 0 <= slot < self.slots:
            raise ValueError('no such slot')
        off = slot * self.slot_bytes + self.rank * self.block_bytes
        order = [(self.rank + 1 + i) % self.world for i in range(self.world)]
        vm = [self.bases[q] + off for q in order]
        return vm, [v + self.vmag_bytes for v in vm]

    def own_destination(self, slot: int):
        """This rank's block of ``slot`` in its OWN buffer only (then ``push_own`` it)."""
        vm, ls = self.destinations(slot)
        return vm[-1:], ls[-1:]

    def vmag_all(self, slot: int):
        """``[world, n_vrows, ld]`` float64 view of this rank's copy of ``slot`` (no copy)."""
        dev = self._torch.device('cuda', self.device)
        a = _DevArray(self._local + slot * self.slot_bytes, (self.world, self.n_vrows, self.ld), '<f8',
                      (self.block_bytes, self.ld * 8, 8))
        return self._torch.as_tensor(a, device=dev)

    def loss_all(self, slot: int):
        """``[world, ld]`` complex128 view of this rank's copy of ``slot`` (no copy)."""
        dev = self._torch.device('cuda', self.device)
        a = _DevArray(self._local + slot * self.slot_bytes + self.vmag_bytes, (self.world, self.ld), '<c16',
                      (self.block_bytes, 16))
        return self._torch.as_tensor(a, device=dev)

    # ---- copy-engine transport -----------------------------------------------------------
    def push(self, slot: int, Vmag_rows, loss, stream=None):
        """All-gather step for results that already sit in device memory (the streaming and NR3
        kernels): DMA ``Vmag_rows [n_vrows, ld]`` and ``loss [ld]`` into block ``rank`` of ``slot``
        on every rank, asynchronously on ``stream``."""
        C, _cabi, torch = self._C, self._cabi, self._torch
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        vm, ls = self.destinations(slot)
        n = len(vm)
        arr_v, arr_l = (C.c_void_p * n)(*vm), (C.c_void_p * n)(*ls)
        _cabi.check(_cabi.gp_peer_push(self.device, n, arr_v, Vmag_rows.data_ptr(), self.n_vrows * self.ld * 8,
                                       st.cuda_stream))
        _cabi.check(_cabi.gp_peer_push(self.device, n, arr_l, loss.data_ptr(), self.ld * 16, st.cuda_stream))

    def push_own(self, slot: int, stream=None):
        """The same step when the solve has written this rank's block into its own buffer: one DMA of
        the whole block (|V| rows and losses are contiguous) to every peer."""
        C, _cabi, torch = self._C, self._cabi, self._torch
        st = torch.cuda.current_stream(self.device) if stream is None else stream
        vm, _ = self.destinations(slot)
        if len(vm) > 1:
            arr = (C.c_void_p * (len(vm) - 1))(*vm[:-1])
            _cabi.check(_cabi.gp_peer_push(self.device, len(vm) - 1, arr, vm[-1], self.block_bytes, st.cuda_stream))

    def wait(self, stream=None):
        """After this returns every rank's contribution issued before its own ``wait`` is in place."""
        import torch.distributed as dist
        (self._torch.cuda.current_stream(self.device) if stream is None else stream).synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)

    def close(self):
        if getattr(self, '_local', None) is None:
            return
        import torch.distributed as dist
        self._torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)             # nobody is still writing into a buffer about to go
        for r, b in enumerate(self.bases):
            if r != self.rank and b is not None:
                self._cabi.gp_peer_close(self.device, b)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)             # every mapping is closed before the memory is freed
        self._cabi.gp_peer_free(self.device, self._local)
        self._local = None

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

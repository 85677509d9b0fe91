"""Multi-GPU partitioning of the path: independent units, no exchange step.

The reference's only parallelism is nn.DataParallel (sff_scripts_interp/main_ms.py:97-103).
Here: one process per GPU (torchrun), the restoration targets of a stack
(target k is interpolated from sections k-1 and k+1,
sff_scripts_interp/inference.py:69-70) or the samples of a batch are split
contiguously over ranks, every rank runs the kernels on its own units with no
communication, and the only collective is one gather of the restored sections
(NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `n_units` for `rank`; sizes differ by at most one."""
    if n_units < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_units, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stack_targets(n_sections: int) -> List[Tuple[int, int, int]]:
    """(k-1, k, k+1) for every interior section k of a stack."""
    return [(k - 1, k, k + 1) for k in range(1, n_sections - 1)]


def max_units_per_rank(n_units: int, world_size: int) -> int:
    return -(-n_units // world_size)


def gather_sections(local: torch.Tensor, n_units: int, group=None, dst=None):
    """Gather the per-rank outputs [n_local, ...] into [n_units, ...] in unit order.

    ``dst=None``: every rank gets the result (one all_gather_into_tensor).  ``dst=r``: only the rank
    whose rank INSIDE ``group`` is ``r`` does (group-local, like ``rank``; translated to the global rank
    torch.distributed.gather expects) (what nn.DataParallel's output gather does in the reference, main_ms.py:97-103);
    the other ranks return None -- 1/world_size of the traffic.
    Ranks may hold different counts (n_units % world_size != 0): each rank pads to
    the maximum, padding is dropped on arrival.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return local
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)                          # group-local rank; `dst` is group-local too
    lo, hi = shard_range(n_units, rank, ws)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} units, expected {hi - lo}")
    m = max_units_per_rank(n_units, ws)
    padded = local
    if local.shape[0] < m:
        pad = torch.zeros((m - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], 0)
    if dst is not None:
        bufs = out = None
        if rank == dst:
            # receive straight into slices of the result: no per-rank staging buffers, no concatenation pass
            # (on 8 GPUs those copies cost rank 0 -- and through the collective everybody -- 8 % of the step)
            out = torch.empty((ws * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            bufs = list(out.split(m, 0))
        gdst = dist.get_global_rank(group, dst) if group is not None else dst
        dist.gather(padded.contiguous(), bufs, dst=gdst, group=group)
        if rank != dst:
            return None
    else:
        out = torch.empty((ws * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if n_units == ws * m:                                # every rank holds m units: `out` already is the answer
        return out
    pieces = []
    for r in range(ws):
        rlo, rhi = shard_range(n_units, r, ws)
        pieces.append(out[r * m: r * m + (rhi - rlo)])
    return torch.cat(pieces, 0)

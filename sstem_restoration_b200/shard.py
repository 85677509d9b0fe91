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


class SectionGatherer:
    """Repeated gathers of equally shaped per-rank outputs to one rank WITHOUT a collective kernel.

    ``dist.gather`` is send / recv kernels that occupy SMs next to kernels written for exactly 2 CTAs per SM and 255
    registers -- measured 3.4 % of the training step at 8 GPUs.  Here the destination buffer ``[world * m, ...]`` is
    symmetric memory (``torch.distributed._symmetric_memory``: every rank maps rank ``dst``'s copy over NVLink); a
    gather is then ONE peer-to-peer ``copy_`` of the local shard into its slice on the copy engine (no SM is touched),
    followed by a stream-ordered barrier on the signal pads so that ``dst`` knows every slice has landed.

    Falls back to :func:`gather_sections` (NCCL / gloo gather) when symmetric memory is unavailable -- CPU tensors, a
    gloo group, an old driver -- and says so in ``mode``.  Every rank must hold ``units_per_rank`` units (pad the
    job, as bench.py's batch shards do).
    """

    def __init__(self, unit_shape, dtype, units_per_rank: int, device, group=None, dst: int = 0, force_collective: bool = False):
        self.group = group
        self.dst = dst
        self.m = int(units_per_rank)
        self.unit_shape = tuple(unit_shape)
        self.dtype = dtype
        self.mode = "collective"
        self.why = None
        self.buf = self.hdl = self.remote = None
        self.flip = 0
        if not (dist.is_available() and dist.is_initialized()):
            self.ws, self.rank, self.mode = 1, 0, "single"
            return
        self.ws = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        device = torch.device(device)
        if force_collective or device.type != "cuda":
            self.why = "forced" if force_collective else "not a CUDA device"
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem
            pg = group if group is not None else dist.group.WORLD
            # two buffers alternate, so a gather may overlap the consumer of the previous one
            self.buf = symm_mem.empty((2, self.ws * self.m) + self.unit_shape, dtype=dtype, device=device)
            self.hdl = symm_mem.rendezvous(self.buf, pg)
            self.remote = self.hdl.get_buffer(dst, tuple(self.buf.shape), dtype)
            self.mode = "p2p-copy-engine"
        except Exception as exc:                           # noqa: BLE001 -- any failure means: use the collective
            self.buf = self.hdl = self.remote = None
            self.why = f"{type(exc).__name__}: {exc}"[:200]

    def gather(self, local: torch.Tensor):
        """local [units_per_rank, *unit_shape] -> [world * units_per_rank, *unit_shape] on ``dst`` (a view of the
        symmetric buffer, valid until the gather after next), None elsewhere.  Stream-ordered on the current stream."""
        if tuple(local.shape) != (self.m,) + self.unit_shape:
            raise ValueError(f"SectionGatherer: expected a shard of shape {(self.m,) + self.unit_shape}, got {tuple(local.shape)}")
        if self.mode == "single":
            return local
        if self.mode != "p2p-copy-engine":
            return gather_sections(local, self.ws * self.m, group=self.group, dst=self.dst)
        k = self.flip
        self.flip ^= 1
        self.remote[k, self.rank * self.m:(self.rank + 1) * self.m].copy_(local, non_blocking=True)   # peer write, copy engine
        self.hdl.barrier(channel=k)                        # signal-pad barrier, ordered after the copy on this stream
        return self.buf[k] if self.rank == self.dst else None

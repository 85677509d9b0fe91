"""Stack restoration loop on the device: BASELINE config 5.

The reference restores a stack one target at a time (sff_scripts_interp/inference.py:69-89): section k is
interpolated from sections k-1 and k+1 -- PNG -> ``/255`` -> x3 replicate -> H2D -> KPN -> sepconv tail -> ``.cpu()``
-> ``*255`` -> uint8 -- and the correction module (sff_scripts_fusion/inference.py:125-171) then warps the degraded
section k with a predicted flow and stitches it with the interpolated one.  The networks that predict the taps and
the flow are out of scope here: they enter as callables.  Everything around them is this module:

  * only uint8 sections cross PCIe (1 byte per pixel each way); every section a rank needs is uploaded once, on a
    copy stream, a few targets ahead of the kernels;
  * ``/255``, x3 replicate (:func:`stack_io.sections_to_input`), the two sepconvs + add + channel mean
    (:func:`sepconv.interpolation_tail`, one launch), ``*255`` -> uint8 (:func:`stack_io.prediction_to_uint8`),
    the flow warp (:class:`warp.SpatialTransformation`) and the stitch mask (:func:`warp_stitch`) run as kernels;
  * targets are split contiguously over ranks (:func:`shard.shard_range`), ranks never talk while computing, and the
    restored uint8 sections are gathered once at the end (:func:`shard.gather_sections`).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from . import _lib, shard
from .sepconv import interpolation_tail, interpolation_tail_tiled
from .stack_io import prediction_to_uint8, sections_to_input
from .warp import SpatialTransformation


def warp_stitch(warped: torch.Tensor, interp_u8: torch.Tensor, want_gray: bool = True, out=None):
    """sff_scripts_fusion/inference.py:163-171 on the device.

    ``warped``: float32 CUDA ``[B,C,H,W]`` (C = 1 or 3), the warped degraded section; ``interp_u8``: uint8 CUDA
    ``[B,H,W]``, the interpolated section.  Returns ``(warped_gray_u8, stitch_u8)``, both uint8 ``[B,H,W]``:
    ``warped_gray = PIL 'L' of (warped*255).astype(uint8)``, ``stitch = where(warped_gray >= 2, warped_gray, interp)``.
    ``out``: optional ``(gray, stitch)`` contiguous uint8 CUDA tensors ``[B,H,W]`` to write into.
    """
    if not (warped.is_cuda and interp_u8.is_cuda):
        raise _lib.SstemError("warp_stitch: CUDA tensors required; there is no CPU fallback")
    if warped.dtype != torch.float32 or interp_u8.dtype != torch.uint8:
        raise TypeError("warp_stitch: float32 warped section and uint8 interpolated section required")
    B, C, H, W = warped.shape
    if tuple(interp_u8.shape) != (B, H, W):
        raise ValueError("warp_stitch: interp_u8 must be [B,H,W] matching warped [B,C,H,W]")
    warped, interp_u8 = warped.contiguous(), interp_u8.contiguous()
    if out is not None:
        gray, stitch = out
        for t in (gray, stitch):
            if t is not None and not (t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous() and tuple(t.shape) == (B, H, W)):
                raise ValueError("warp_stitch: out tensors must be contiguous uint8 CUDA [B,H,W]")
        want_gray = gray is not None
    else:
        gray = torch.empty((B, H, W), dtype=torch.uint8, device=warped.device) if want_gray else None
        stitch = torch.empty((B, H, W), dtype=torch.uint8, device=warped.device)
    if stitch.numel():
        code = _lib.load().sstem_warp_stitch_u8(warped.data_ptr(), interp_u8.data_ptr(), gray.data_ptr() if want_gray else None,
                                                stitch.data_ptr(), B, C, H, W, torch.cuda.current_stream(warped.device).cuda_stream)
        if code:
            _lib.check(code, "sstem_warp_stitch_u8")
    return gray, stitch


def warp_and_stitch(moving: torch.Tensor, flow: torch.Tensor, interp_u8: torch.Tensor, want_gray: bool = True, out=None):
    """``SpatialTransformation(moving, flow)`` followed by :func:`warp_stitch`, as ONE kernel: the stitch assembly is the warp
    kernel's epilogue, so the float32 warped image is never written or read back (sff_scripts_fusion/inference.py:150 +
    :163-171).  ``moving`` [B,C,H,W] float32 CUDA (C = 1 or 3), ``flow`` [B,H,W,2] (any strides), ``interp_u8`` [B,H,W].
    Returns ``(warped_gray_u8, stitch_u8)``; bit-equal to the two separate calls."""
    if not (moving.is_cuda and flow.is_cuda and interp_u8.is_cuda):
        raise _lib.SstemError("warp_and_stitch: CUDA tensors required; there is no CPU fallback")
    if moving.dtype != torch.float32 or flow.dtype != torch.float32 or interp_u8.dtype != torch.uint8:
        raise TypeError("warp_and_stitch: float32 image / flow and uint8 interpolated section required")
    B, C, H, W = moving.shape
    if tuple(flow.shape) != (B, H, W, 2) or tuple(interp_u8.shape) != (B, H, W):
        raise ValueError("warp_and_stitch: flow must be [B,H,W,2] and interp_u8 [B,H,W] matching moving [B,C,H,W]")
    moving, interp_u8 = moving.contiguous(), interp_u8.contiguous()
    if out is not None:
        gray, stitch = out
        want_gray = gray is not None
    else:
        gray = torch.empty((B, H, W), dtype=torch.uint8, device=moving.device) if want_gray else None
        stitch = torch.empty((B, H, W), dtype=torch.uint8, device=moving.device)
    if stitch.numel():
        import ctypes
        strides = (ctypes.c_int64 * 4)(*flow.stride())
        code = _lib.load().sstem_warp_stitch_forward(moving.data_ptr(), flow.data_ptr(), strides, interp_u8.data_ptr(),
                                                     gray.data_ptr() if want_gray else None, stitch.data_ptr(), B, C, H, W,
                                                     torch.cuda.current_stream(moving.device).cuda_stream)
        if code:
            _lib.check(code, "sstem_warp_stitch_forward")
    return gray, stitch


class _SectionCache:
    """uint8 sections of this rank on the device: each is uploaded once, on a copy stream, ahead of its first use."""

    def __init__(self, stack, device, depth):
        self.stack, self.device, self.depth = stack, device, depth
        self.on_device = stack.is_cuda
        self.copy_stream = None if self.on_device else torch.cuda.Stream(device=device)
        self.slots: Dict[int, tuple] = {}
        self.h2d_bytes = 0

    def request(self, k):
        if self.on_device or k in self.slots or not (0 <= k < self.stack.shape[0]):
            return
        with torch.cuda.stream(self.copy_stream):
            t = self.stack[k].to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.slots[k] = (t, ev)
        self.h2d_bytes += t.numel()

    def get(self, k):
        if self.on_device:
            return self.stack[k]
        self.request(k)
        t, ev = self.slots[k]
        torch.cuda.current_stream(self.device).wait_event(ev)
        return t

    def drop_below(self, k):
        for j in [j for j in self.slots if j < k]:
            t, _ = self.slots.pop(j)
            t.record_stream(torch.cuda.current_stream(self.device))


def restore_stack(stack: torch.Tensor, taps_fn: Callable, flow_fn: Optional[Callable] = None, *,
                  rank: int = 0, world_size: int = 1, group=None, dst: Optional[int] = 0, device=None,
                  prefetch: int = 3, to_host: bool = False, host_out: Optional[dict] = None):
    """Restore the interior sections of ``stack`` (uint8 ``[N,H,W]``, pinned host memory or CUDA).

    For every target k in 1..N-2 owned by this rank::

        x = sections_to_input(stack[k-1], stack[k+1])                # [1,6,H,W] float32, gray x3, /255
        k1v, k1h, k2v, k2h = taps_fn(k, x)                           # the KPN (out of scope): four [1,51,H,W] tensors, or tile-major
        interp = prediction_to_uint8(interpolation_tail(x[:, :3], x[:, 3:6], k1v, k1h, k2v, k2h))
        if flow_fn:                                                  # the correction module's flow net (out of scope)
            xk = sections_to_input(stack[k])                         # [1,3,H,W]: input_sff
            warped_gray, stitch = warp_and_stitch(xk[:, :1], flow_fn(k, xk, interp), interp)   # one kernel; planes identical

    Returns a dict of uint8 tensors -- ``interp`` and, with ``flow_fn``, ``warped`` and ``stitch`` -- plus ``stats``.

    ``to_host=False`` (default): device tensors ``[N-2,H,W]`` gathered to rank ``dst`` (every rank when ``dst`` is None;
    other ranks get ``None``) -- the path's only collective.
    ``to_host=True``: every rank downloads ITS OWN targets (what a rank would hand to its PNG writers) into pinned host
    tensors ``[n_local,H,W]`` -- ``host_out`` may supply them -- one asynchronous copy per target on a download stream,
    overlapped with the next target's kernels; no collective; the call returns when the copies have completed.
    """
    if not torch.cuda.is_available():
        raise _lib.SstemError("restore_stack: no CUDA device; there is no CPU fallback")
    if stack.dtype != torch.uint8 or stack.dim() != 3 or stack.shape[0] < 3:
        raise ValueError("restore_stack: stack must be uint8 [N,H,W] with N >= 3")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if stack.is_cuda and stack.device != dev:
        raise ValueError("restore_stack: a CUDA stack must live on the computing device")
    N, H, W = stack.shape
    targets = shard.stack_targets(N)
    lo, hi = shard.shard_range(len(targets), rank, world_size)
    mine = targets[lo:hi]
    n0 = _lib.launch_count()
    cache = _SectionCache(stack, dev, prefetch)
    warp = SpatialTransformation(True)
    names = ("interp",) + (("warped", "stitch") if flow_fn is not None else ())
    local = {n: torch.empty((len(mine), H, W), dtype=torch.uint8, device=dev) for n in names}
    host = None
    if to_host:
        host = host_out if host_out is not None else {n: torch.empty((len(mine), H, W), dtype=torch.uint8).pin_memory() for n in names}
        for n in names:
            if tuple(host[n].shape) != (len(mine), H, W) or host[n].dtype != torch.uint8 or host[n].is_cuda:
                raise ValueError(f"restore_stack: host_out[{n!r}] must be a CPU uint8 tensor [{len(mine)},{H},{W}]")
        down = torch.cuda.Stream(device=dev)
    with torch.cuda.device(dev), torch.no_grad():
        for j in range(min(prefetch, len(mine))):
            for k in mine[j]:
                cache.request(k)
        for i, (ka, k, kb) in enumerate(mine):
            if i + prefetch < len(mine):
                for kk in mine[i + prefetch]:
                    cache.request(kk)
            x = sections_to_input(cache.get(ka), cache.get(kb), 0)
            k1v, k1h, k2v, k2h = taps_fn(k, x)
            # [1,51,H,W] taps: the fused tail; tile-major taps (6-D, what ModuleTapProducer writes): the tile-major tail
            tail = interpolation_tail_tiled if k1v.dim() == 6 else interpolation_tail
            interp = prediction_to_uint8(tail(x[:, :3], x[:, 3:6], k1v, k1h, k2v, k2h), 0, out=local["interp"][i:i + 1])
            if flow_fn is not None:
                xk = sections_to_input(cache.get(k), None, 0)
                # input_sff is a gray section replicated x3 (inference.py:129-131): its three warped planes are identical and
                # PIL's 'L' of an R = G = B triple is R, so one plane is warped and stitched (a third of the bytes, same bits)
                warp_and_stitch(xk[:, :1].contiguous(), flow_fn(k, xk, interp), interp,
                                out=(local["warped"][i:i + 1], local["stitch"][i:i + 1]))
            if to_host:
                done = torch.cuda.Event()
                done.record()
                with torch.cuda.stream(down):
                    down.wait_event(done)
                    for n in names:
                        host[n][i].copy_(local[n][i], non_blocking=True)
            cache.drop_below(ka)
        out = {}
        if to_host:
            down.synchronize()
            out.update({n: host[n] for n in names})
        else:
            for n in names:
                out[n] = shard.gather_sections(local[n], len(targets), group=group, dst=dst) if world_size > 1 else local[n]
    out["stats"] = {"targets": len(targets), "targets_this_rank": len(mine), "h2d_bytes": cache.h2d_bytes,
                    "d2h_bytes": sum(int(host[n].numel()) for n in names) if to_host else 0,
                    "kernel_launches": _lib.launch_count() - n0}
    return out

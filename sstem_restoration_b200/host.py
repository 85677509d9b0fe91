"""Host-buffer entry point: the sepconv forward + tap gradients for tensors that live in
(pinned) host memory, as they do when the caller is a CPU data pipeline.

The batch is cut into chunks; each chunk's host->device copies, the two sm_100a kernels
(through the C ABI) and the device->host copies of its results are queued on one of a few
CUDA streams, so chunk i+1 uploads while chunk i computes and chunk i-1 downloads -- the
PCIe link runs full duplex and the kernels disappear behind it.  Semantics are those of
``SeparableConvolution.apply`` followed by ``.backward(grad_output)``
(libs/sepconv/SeparableConvolution.py:16-78 of the reference).
"""
from __future__ import annotations

import threading
from typing import Optional, Tuple

import torch

from . import _lib


class _Workspace:
    """Device staging buffers for one stream, grown on demand and reused across calls."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, shape):
        n = 1
        for s in shape:
            n *= int(s)
        buf = self.bufs.get(name)
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, dtype=torch.float32, device=self.device)
            self.bufs[name] = buf
        return buf[:n].view(*shape)


_streams = {}
_workspaces = {}
_res_lock = threading.Lock()


def _resources(device, n):
    """Side streams + staging buffers of (calling thread, device): replica threads (nn.DataParallel) never share
    a workspace, so two concurrent calls cannot overwrite each other's staging buffers."""
    key = (threading.get_ident(), device.index, n)
    with _res_lock:
        if key not in _streams:
            _streams[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
            _workspaces[key] = [_Workspace(device) for _ in range(n)]
        return _streams[key], _workspaces[key]


def sepconv_forward_backward_host(input: torch.Tensor, vertical: torch.Tensor, horizontal: torch.Tensor,
                                  grad_output: Optional[torch.Tensor] = None, *, device=None, chunk: int = 2,
                                  n_streams: int = 3,
                                  out: Optional[Tuple[torch.Tensor, ...]] = None, join: bool = True):
    """input [B,C,H+50,W+50], vertical/horizontal [B,51,H,W], grad_output [B,C,H,W]: CPU float32,
    ideally pinned.  Returns CPU (pinned) ``output`` or ``(output, grad_vertical, grad_horizontal)``.
    ``out`` may supply the pinned result tensors to reuse.  With ``join=True`` (default) the call returns
    after the last device->host copy has COMPLETED (host-side synchronisation of the side streams): the
    returned CPU tensors can be read immediately.  With ``join=False`` the call only queues the work
    (back-to-back calls then overlap one call's downloads with the next call's uploads); the CPU results
    are NOT valid until :func:`join_host_pipeline` (or a device synchronize) has returned."""
    if not torch.cuda.is_available():
        raise _lib.SstemError("sepconv_forward_backward_host: no CUDA device; there is no CPU fallback")
    for t in (input, vertical, horizontal):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError("host API expects contiguous float32 CPU tensors")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    B, C = input.shape[:2]
    K, H, W = vertical.shape[1:]
    assert input.shape[2] - K == H - 1 and input.shape[3] - K == W - 1 and vertical.shape == horizontal.shape
    want_grad = grad_output is not None
    if out is None:
        res = [torch.empty((B, C, H, W), dtype=torch.float32).pin_memory()]
        if want_grad:
            res += [torch.empty_like(vertical).pin_memory(), torch.empty_like(horizontal).pin_memory()]
    else:
        res = list(out)
    lib = _lib.load()
    streams, spaces = _resources(dev, n_streams)
    with torch.cuda.device(dev):
        for ci, lo in enumerate(range(0, B, chunk)):
            hi = min(B, lo + chunk)
            n = hi - lo
            s, ws = streams[ci % n_streams], spaces[ci % n_streams]
            with torch.cuda.stream(s):
                d_in = ws.get("in", (n,) + tuple(input.shape[1:]))
                d_v = ws.get("v", (n, K, H, W))
                d_h = ws.get("h", (n, K, H, W))
                d_out = ws.get("out", (n, C, H, W))
                # all uploads of the chunk first, then the kernels, then all downloads: the copy engines
                # serve their queues in issue order, so a copy that waits on a kernel must not sit in
                # front of the next chunk's uploads
                d_in.copy_(input[lo:hi], non_blocking=True)
                d_v.copy_(vertical[lo:hi], non_blocking=True)
                d_h.copy_(horizontal[lo:hi], non_blocking=True)
                if want_grad:
                    d_g = ws.get("g", (n, C, H, W))
                    d_gv = ws.get("gv", (n, K, H, W))
                    d_gh = ws.get("gh", (n, K, H, W))
                    d_g.copy_(grad_output[lo:hi], non_blocking=True)
                _lib.check(lib.sstem_sepconv_forward(d_in.data_ptr(), d_v.data_ptr(), d_h.data_ptr(), d_out.data_ptr(),
                                                     n, C, H, W, K, 0, s.cuda_stream), "sstem_sepconv_forward")
                if want_grad:
                    _lib.check(lib.sstem_sepconv_backward(d_g.data_ptr(), d_in.data_ptr(), d_v.data_ptr(), d_h.data_ptr(),
                                                          None, d_gv.data_ptr(), d_gh.data_ptr(), n, C, H, W, K, 0,
                                                          s.cuda_stream), "sstem_sepconv_backward")
                res[0][lo:hi].copy_(d_out, non_blocking=True)
                if want_grad:
                    res[1][lo:hi].copy_(d_gv, non_blocking=True)
                    res[2][lo:hi].copy_(d_gh, non_blocking=True)
    if join:
        join_host_pipeline(dev, n_streams)
    return tuple(res) if want_grad else res[0]


def join_host_pipeline(device=None, n_streams: int = 3, host_sync: bool = True) -> None:
    """Wait for everything this thread queued with sepconv_forward_backward_host on `device`.

    ``host_sync=True`` (default): block the calling host thread until the side streams have drained -- the
    pinned CPU results are then complete and safe to read.  ``host_sync=False``: only make the current CUDA
    stream wait for the side streams (a device-side dependency; the host is NOT blocked and CPU results must
    not be read yet) -- for callers that keep queueing device work behind the results."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    streams, _ = _resources(dev, n_streams)
    cur = torch.cuda.current_stream(dev)
    for s in streams:
        cur.wait_stream(s)
        if host_sync:
            s.synchronize()

"""Stack pre/post-processing around the interpolation path, on the device (SURVEY.md section
8f, N4): only uint8 sections cross PCIe.

  * :func:`sections_to_input` -- sff_scripts_interp/inference.py:69-83: two uint8 sections ->
    the network's float32 input ``[B,6,H+2*PAD,W+2*PAD]`` (x3 replicate, /255, zero pad);
  * :func:`prediction_to_uint8` -- inference.py:84-88: ``(F.pad(pred, (-PAD,)*4) * 255).astype(uint8)``.

Both are bit-equal to the numpy expressions.  CUDA tensors in -> CUDA tensors out; numpy / CPU
tensors are uploaded (pinned staging is the caller's business) and the result stays on the device
for :func:`sections_to_input` (its consumer is the network) and is downloaded for
:func:`prediction_to_uint8`.  No CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib


def _cuda(t, dtype):
    if not torch.cuda.is_available():
        raise _lib.SstemError("sstem_restoration_b200 stack_io: no CUDA device; there is no CPU fallback")
    t = torch.as_tensor(t)
    if t.dtype != dtype:
        raise TypeError(f"stack_io: expected {dtype}, got {t.dtype}")
    host = not t.is_cuda
    if host:
        t = t.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
    return t.contiguous(), host


def sections_to_input(section_prev, section_next=None, pad=0):
    """uint8 ``[H,W]`` or ``[B,H,W]`` sections k-1 and k+1 -> float32 ``[B,6,H+2*pad,W+2*pad]`` on the device.
    ``section_next=None``: one section -> ``[B,3,H+2*pad,W+2*pad]`` (gray x3, /255: the correction module's
    ``input_sff``, sff_scripts_fusion/inference.py:127-131)."""
    a, _ = _cuda(section_prev, torch.uint8)
    b = None
    if section_next is not None:
        b, _ = _cuda(section_next, torch.uint8)
    if a.dim() == 2:
        a = a[None]
        b = b[None] if b is not None else None
    if a.dim() != 3 or (b is not None and (a.shape != b.shape or a.device != b.device)):
        raise ValueError("sections_to_input: two uint8 sections of the same shape [H,W] or [B,H,W]")
    B, H, W = a.shape
    out = torch.empty((B, 6 if b is not None else 3, H + 2 * pad, W + 2 * pad), dtype=torch.float32, device=a.device)
    if out.numel():
        code = _lib.load().sstem_sections_to_input(a.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(), B, H, W, pad,
                                                   torch.cuda.current_stream(a.device).cuda_stream)
        if code:
            _lib.check(code, "sstem_sections_to_input")
    return out


def prediction_to_uint8(pred, pad=0, out=None):
    """float32 ``[B,1,H+2*pad,W+2*pad]`` (or ``[H+2*pad,W+2*pad]``) -> uint8 ``[B,H,W]`` (``[H,W]``).
    ``out``: optional contiguous uint8 CUDA tensor ``[B,H,W]`` to write into (e.g. a slice of a stack)."""
    p, host = _cuda(pred, torch.float32)
    squeeze = p.dim() == 2
    if squeeze:
        p = p[None, None]
    if p.dim() != 4 or p.size(1) != 1:
        raise ValueError("prediction_to_uint8: pred must be [B,1,H,W] or [H,W]")
    B, _, OH, OW = p.shape
    H, W = OH - 2 * pad, OW - 2 * pad
    if H <= 0 or W <= 0:
        raise ValueError("prediction_to_uint8: pad larger than the prediction")
    if out is None:
        out = torch.empty((B, H, W), dtype=torch.uint8, device=p.device)
    elif not (out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == (B, H, W) and out.device == p.device):
        raise ValueError("prediction_to_uint8: out must be a contiguous uint8 CUDA tensor [B,H,W] on pred's device")
    code = _lib.load().sstem_prediction_to_u8(p.data_ptr(), out.data_ptr(), B, H, W, pad,
                                              torch.cuda.current_stream(p.device).cuda_stream)
    if code:
        _lib.check(code, "sstem_prediction_to_u8")
    if squeeze:
        out = out[0]
    return out.cpu().numpy() if host else out

"""Flow-driven bilinear backward warps: the reference's call signatures over the
sm_100a C ABI.

  * ``SpatialTransformation(use_gpu)(moving_image[B,C,H,W], deformation_matrix[B,H,W,2])``
    mirrors sff_scripts_unfolding/utils/image_warp_torch.py:5-113 (identical copy in
    sff_scripts_fusion/utils/).  One CUDA kernel replaces ~30 ATen ops, the per-call
    CPU meshgrid / int64 base-index construction and their two H2D copies (:11-29,61).
    Results are bit-equal to the reference run on CPU.
  * ``image_warp(im, flow, mode)`` mirrors numpy simu_sff/image_warp.py:3-111
    (clamp border with the x1-from-clipped-x0 quirk, uint8 truncation).

Host buffers (CPU tensors / numpy arrays) are accepted: they are copied to the
current CUDA device, warped by the kernel and copied back -- still the CUDA
path.  Without the compiled library or without a GPU every call raises.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def _stream_ptr(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.SstemError("sstem_restoration_b200 warp: no CUDA device; there is no CPU fallback")


def _warp_launch(moving, flow, nhwc_memory):
    """moving [B,C,H,W] CUDA float32 contiguous, flow [B,H,W,2] CUDA float32 (any strides)."""
    B, C, H, W = moving.shape
    if nhwc_memory:
        out = torch.empty((B, H, W, C), dtype=torch.float32, device=moving.device)
    else:
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=moving.device)
    if out.numel() == 0:
        return out.permute(0, 3, 1, 2) if nhwc_memory else out
    strides = (ctypes.c_int64 * 4)(*flow.stride())
    # the library launches on the device that owns `out`; only the stream has to be the caller's
    code = _lib.load().sstem_warp_forward(
        moving.data_ptr(), flow.data_ptr(), strides, out.data_ptr(), B, C, H, W,
        _lib.LAYOUT_NHWC if nhwc_memory else _lib.LAYOUT_NCHW, _stream_ptr(moving.device))
    if code:
        _lib.check(code, "sstem_warp_forward")
    return out.permute(0, 3, 1, 2) if nhwc_memory else out


class _WarpFunction(torch.autograd.Function):
    """Differentiable like the reference's module (its ATen ops carry autograd: image_warp_torch.py:32-95): gradients
    w.r.t. the moving image (scatter of the tap weights) and the flow, through ``sstem_warp_backward``."""

    @staticmethod
    def forward(ctx, moving, flow, nhwc_memory):
        ctx.save_for_backward(moving, flow)
        return _warp_launch(moving, flow, nhwc_memory)

    @staticmethod
    def backward(ctx, grad):
        moving, flow = ctx.saved_tensors
        need_m, need_f = ctx.needs_input_grad[:2]
        if not (need_m or need_f):
            return None, None, None
        B, C, H, W = moving.shape
        if grad.is_cuda == False:
            raise NotImplementedError()
        grad = grad.to(torch.float32).contiguous()             # [B,C,H,W] (a permuted NHWC view becomes NCHW here)
        gm = torch.empty_like(moving) if need_m else None
        gf = torch.empty((B, H, W, 2), dtype=torch.float32, device=moving.device) if need_f else None
        if moving.numel():
            strides = (ctypes.c_int64 * 4)(*flow.stride())
            code = _lib.load().sstem_warp_backward(
                moving.data_ptr(), flow.data_ptr(), strides, grad.data_ptr(), gm.data_ptr() if need_m else None,
                gf.data_ptr() if need_f else None, B, C, H, W, _stream_ptr(moving.device))
            if code:
                _lib.check(code, "sstem_warp_backward")
        return gm, gf, None


class SpatialTransformation(nn.Module):
    """Drop-in for image_warp_torch.SpatialTransformation.

    ``use_gpu`` is kept for signature compatibility; the computation always runs
    on a CUDA device (the tensor's own, or the current one for host tensors).
    ``nhwc_memory=True`` reproduces the reference's memory format (a permuted view
    of NHWC storage, image_warp_torch.py:94,112); the default returns contiguous
    NCHW with identical values.
    """

    def __init__(self, use_gpu=False, nhwc_memory=False):
        self.use_gpu = use_gpu
        super(SpatialTransformation, self).__init__()
        self.nhwc_memory = nhwc_memory

    def forward(self, moving_image, deformation_matrix):
        _require_cuda()
        if moving_image.dim() != 4 or deformation_matrix.dim() != 4 or deformation_matrix.size(3) != 2:
            raise ValueError("expected moving_image [B,C,H,W] and deformation_matrix [B,H,W,2]")
        B, C, H, W = moving_image.shape
        if tuple(deformation_matrix.shape[:3]) != (B, H, W):
            raise ValueError("deformation_matrix must be [B,H,W,2] matching moving_image")
        host = not moving_image.is_cuda
        if host or moving_image.dtype != torch.float32 or deformation_matrix.dtype != torch.float32 \
                or deformation_matrix.device != moving_image.device:
            dev = torch.device("cuda", torch.cuda.current_device()) if host else moving_image.device
            moving_image = moving_image.to(device=dev, dtype=torch.float32, non_blocking=True)
            deformation_matrix = deformation_matrix.to(device=dev, dtype=torch.float32, non_blocking=True)  # strides kept
        if not moving_image.is_contiguous():
            moving_image = moving_image.contiguous()
        if torch.is_grad_enabled() and (moving_image.requires_grad or deformation_matrix.requires_grad):
            out = _WarpFunction.apply(moving_image, deformation_matrix, self.nhwc_memory)
        else:
            out = _warp_launch(moving_image, deformation_matrix, self.nhwc_memory)
        return out.cpu() if host else out


def image_warp(im, flow, mode="bilinear"):
    """Drop-in for numpy ``image_warp`` (simu_sff/image_warp.py:3).

    ``im``: ndim 2/3/4 = [[B],H,W,[C]], uint8 or float32; ``flow``: [[B],H,W,2] float32.
    numpy in -> numpy uint8 out (through the GPU); CUDA tensors in -> CUDA uint8 tensor out.
    """
    _require_cuda()
    is_np = isinstance(im, np.ndarray)
    if is_np:
        dev = torch.device("cuda", torch.cuda.current_device())
        if im.dtype not in (np.uint8, np.float32):
            raise TypeError(f"image_warp: uint8 or float32 image required, got {im.dtype}")
        flow_np = np.asarray(flow)
        if flow_np.dtype != np.float32:
            raise TypeError(f"image_warp: float32 flow required, got {flow_np.dtype}")
        im_t = torch.from_numpy(np.ascontiguousarray(im)).to(dev, non_blocking=True)
        flow_t = torch.from_numpy(np.ascontiguousarray(flow_np)).to(dev, non_blocking=True)
    else:
        if not (torch.is_tensor(im) and im.is_cuda):
            raise TypeError("image_warp: numpy arrays or CUDA tensors required")
        if im.dtype not in (torch.uint8, torch.float32) or flow.dtype != torch.float32:
            raise TypeError("image_warp: uint8/float32 image and float32 flow required")
        im_t, flow_t = im.contiguous(), flow.to(im.device).contiguous()
    out, _ = _image_warp_cuda(im_t, flow_t, mode, want_float=False)
    return out.cpu().numpy() if is_np else out


def _image_warp_cuda(im_t, flow_t, mode="bilinear", want_float=False, want_u8=True):
    """im_t [[B],H,W,[C]] CUDA uint8/float32; returns (uint8, float-before-cast) shaped like im_t."""
    nd = im_t.dim()
    if nd == 2:
        im4, fl4 = im_t[None, :, :, None], flow_t[None]
    elif nd == 3:
        im4, fl4 = im_t[None], flow_t[None]
    elif nd == 4:
        im4, fl4 = im_t, flow_t
    else:
        raise AttributeError("The dimension of im must be 2, 3 or 4")
    if mode == "bilinear":
        m = _lib.WARP_BILINEAR
    elif mode == "nearest":
        m = _lib.WARP_NEAREST
    else:
        raise UnboundLocalError("image_warp: mode must be 'nearest' or 'bilinear'")
    B, H, W, C = im4.shape
    if tuple(fl4.shape) != (B, H, W, 2):
        raise ValueError(f"flow shape {tuple(flow_t.shape)} does not match image {tuple(im_t.shape)}")
    im4, fl4 = im4.contiguous(), fl4.contiguous()
    out_u8 = torch.empty((B, H, W, C), dtype=torch.uint8, device=im4.device) if want_u8 else None
    out_f = torch.empty((B, H, W, C), dtype=torch.float32, device=im4.device) if want_float else None
    if im4.numel() > 0:
        code = _lib.load().sstem_image_warp(
            im4.data_ptr(), _lib.PIX_U8 if im4.dtype == torch.uint8 else _lib.PIX_F32, fl4.data_ptr(),
            out_u8.data_ptr() if want_u8 else None, out_f.data_ptr() if want_float else None,
            B, H, W, C, m, _stream_ptr(im4.device))
        if code:
            _lib.check(code, "sstem_image_warp")

    def _shape(t):
        if t is None:
            return None
        if nd == 2:   # image_warp.py:103-104: np.squeeze drops EVERY unit axis of [1,H,W,1]
            return t.reshape([d for d in (H, W) if d != 1])
        return t.reshape(im_t.shape)

    return _shape(out_u8), _shape(out_f)

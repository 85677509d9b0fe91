"""Tap producer: the last two layers of the reference's ``IFNet._kernel_module``

    nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) -> nn.Conv2d(51, 51, 3, 1, 1)

(sff_scripts_interp/model/model_interp.py:18, 130-137; four instances, :34-37, run at :86-89) as one sm_100a kernel
(csrc/tapconv.cu: tcgen05 TF32 implicit GEMM, the upsample folded into the operand producer), writing the taps in the
layout of ``Conv2d`` or directly tile-major for :func:`sstem_restoration_b200.sepconv_forward_tiled`.  Forward only
(the stack-restoration path, c5); no CPU / PyTorch fallback.
"""
import torch

from . import _lib


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def pack_tap_conv_weight(weight: torch.Tensor) -> torch.Tensor:
    """``Conv2d.weight`` ``[cout<=64, cin<=52, 3, 3]`` -> the kernel's resident image (118 KB, TF32-rounded).  Pack once per
    layer, reuse for every call."""
    if weight.is_cuda == False:
        raise NotImplementedError()
    if weight.dtype != torch.float32 or weight.dim() != 4 or tuple(weight.shape[2:]) != (3, 3):
        raise TypeError("pack_tap_conv_weight: float32 [cout, cin, 3, 3] required")
    cout, cin = weight.shape[:2]
    if cin > 52 or cout > 64:
        raise ValueError("pack_tap_conv_weight: cin <= 52 and cout <= 64 required")
    lib = _lib.load()
    packed = torch.empty(lib.sstem_tap_conv3x3_packed_elems(), dtype=torch.float32, device=weight.device)
    code = lib.sstem_tap_conv3x3_pack_weights(weight.contiguous().data_ptr(), packed.data_ptr(), cin, cout, _stream_ptr(weight))
    if code:
        _lib.check(code, "sstem_tap_conv3x3_pack_weights")
    packed._sstem_cin_cout = (int(cin), int(cout))
    return packed


def tap_conv3x3(x: torch.Tensor, packed_weight: torch.Tensor, bias=None, cin=None, cout=None, upsample=True, tiled=False) -> torch.Tensor:
    """``conv2d(upsample2x(x), weight, bias, stride 1, padding 1)`` (``upsample=False``: the convolution alone).

    x ``[B, cin, h, w]`` float32 contiguous; ``packed_weight`` from :func:`pack_tap_conv_weight`.  Returns ``[B, cout, H, W]``
    or, with ``tiled=True`` (cout == 51), the tile-major taps ``[B, ceil(H/8), ceil(W/8), 51, 8, 8]``."""
    if x.is_cuda == False:
        raise NotImplementedError()
    if x.dtype != torch.float32 or x.dim() != 4:
        raise TypeError("tap_conv3x3: float32 [B, cin, h, w] required")
    assert x.is_contiguous() == True
    if torch.is_grad_enabled() and x.requires_grad:
        # forward only: returning a tensor without a graph would silently cut the gradient of everything upstream
        raise NotImplementedError("tap_conv3x3 is forward only (the stack-restoration path): call it under torch.no_grad() or detach its input")
    if cin is None or cout is None:
        cin, cout = getattr(packed_weight, "_sstem_cin_cout", (None, None))
        if cin is None:
            raise ValueError("tap_conv3x3: pass cin / cout (the packed weight does not carry them)")
    B, c, h, w = x.shape
    if c != cin:
        raise ValueError(f"tap_conv3x3: x has {c} channels, the weight {cin}")
    H, W = (2 * h, 2 * w) if upsample else (h, w)
    if tiled:
        if cout != 51:
            raise ValueError("tap_conv3x3: the tile-major layout is defined for 51 taps")
        out = torch.empty((B, (H + 7) // 8, (W + 7) // 8, 51, 8, 8), dtype=torch.float32, device=x.device)
    else:
        out = torch.empty((B, cout, H, W), dtype=torch.float32, device=x.device)
    if out.numel():
        flags = (_lib.TAPCONV_UPSAMPLE2X if upsample else 0) | (_lib.TAPCONV_TILED if tiled else 0)
        code = _lib.load().sstem_tap_conv3x3(x.data_ptr(), packed_weight.data_ptr(), bias.contiguous().data_ptr() if bias is not None else None,
                                             out.data_ptr(), B, cin, cout, h, w, flags, _stream_ptr(x))
        if code:
            _lib.check(code, "sstem_tap_conv3x3")
    return out


class ModuleTapProducer(torch.nn.Module):
    """Drop-in for the tail ``nn.Sequential(upsample, Conv2d(51, 51, 3, 1, 1))`` of ``_kernel_module``: holds ``weight`` /
    ``bias`` under the names of ``nn.Conv2d`` (so the reference layer's state_dict loads), packs the weight on first use
    and whenever it changes."""

    def __init__(self, in_channels=51, out_channels=51, upsample=True, tiled=False):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels, 3, 3))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))
        torch.nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self.upsample, self.tiled = upsample, tiled
        self._packed, self._packed_version = None, None

    def forward(self, x):
        key = (self.weight._version, self.weight.data_ptr())
        if self._packed is None or self._packed_version != key or self._packed.device != x.device:
            self._packed = pack_tap_conv_weight(self.weight.detach())
            self._packed_version = key
        return tap_conv3x3(x, self._packed, self.bias.detach(), upsample=self.upsample, tiled=self.tiled)

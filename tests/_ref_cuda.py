"""ctypes access to oracle/_ref/libref_sepconv.so: the reference's own CUDA source
(libs/sepconv/src/SeparableConvolution_kernel.cu) compiled verbatim for sm_100a by
oracle/Makefile.  Test-only."""
import ctypes
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_sepconv.so")
_lib = None


def available():
    return os.path.exists(REF_SO)


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(REF_SO)
        _lib.ref_sepconv_forward.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_long] * 4 + [ctypes.c_void_p]
        _lib.ref_sepconv_backward.argtypes = [ctypes.c_void_p] * 7 + [ctypes.c_long] * 4 + [ctypes.c_void_p]
    return _lib


def forward(inp, v, h):
    B, C = inp.shape[:2]
    H, W = v.shape[2:]
    out = torch.zeros((B, C, H, W), device=inp.device)
    rc = _load().ref_sepconv_forward(inp.data_ptr(), v.data_ptr(), h.data_ptr(), out.data_ptr(), B, C, H, W,
                                     torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return out


def backward(g, inp, v, h):
    B, C = inp.shape[:2]
    H, W = v.shape[2:]
    gi, gv, gh = torch.zeros_like(inp), torch.zeros_like(v), torch.zeros_like(h)
    rc = _load().ref_sepconv_backward(g.data_ptr(), inp.data_ptr(), v.data_ptr(), h.data_ptr(), gi.data_ptr(),
                                      gv.data_ptr(), gh.data_ptr(), B, C, H, W, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return gi, gv, gh

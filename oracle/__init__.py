"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's hot path (sydeng99/ssTEM-restoration):
the 51-tap adaptive separable local convolution (``libs/sepconv``), the expression
the interpolation network wraps around it (``model_interp.py:90-97``) and the two
flow-driven bilinear backward warps (``image_warp_torch.SpatialTransformation``
and numpy ``image_warp``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package, and only as the checker or as the CPU baseline -- the product package
``sstem_restoration_b200`` never does.

Parity pins (see tests/golden/README.md):
  * warp restatements: bit-checked against the reference's own Python
    (imported from /root/reference in the build container) -- fixtures
    ``tests/golden/warp_*.npz`` made by ``tests/golden/make_golden.py``.
  * sepconv fp32 "reforder" restatement: checked against the reference's own
    ``.cu`` compiled verbatim for sm_100a (``oracle/_ref``) and run on a B200 --
    fixtures ``tests/golden/sepconv_ref_*.npz`` made by
    ``tests/golden/make_sepconv_ref_golden.py``.
  * grad w.r.t. input has no reference implementation (the reference returns
    zeros, ``libs/sepconv/SeparableConvolution.py:60``): its oracle is the
    mathematical adjoint in fp64, cross-checked with torch autograd --
    "parity unpinned" for that one output.

All file:line citations are relative to the reference checkout.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so (and oracle/_ref when the reference is mounted)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "sepconv_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _dims(inp, v):
    B, C, IH, IW = inp.shape
    Bv, K, H, W = v.shape
    assert B == Bv and IH == H + K - 1 and IW == W + K - 1, (inp.shape, v.shape)
    assert K <= 128
    return (ctypes.c_int64(B), ctypes.c_int64(C), ctypes.c_int64(H), ctypes.c_int64(W), ctypes.c_int(K)), (B, C, H, W, K)


# --------------------------------------------------------------------------- sepconv
def sepconv_forward_reforder(inp, v, h):
    """fp32, the reference's summation order (kernel.cu:38-51: fy outer, fx inner,
    one accumulator, FMUL(in,v) then FFMA(.,h,acc))."""
    inp, v, h = _f32(inp), _f32(v), _f32(h)
    d, (B, C, H, W, K) = _dims(inp, v)
    out = np.empty((B, C, H, W), np.float32)
    _lib().oracle_sepconv_fwd_reforder(_p(inp), _p(v), _p(h), _p(out), *d)
    return out


def sepconv_forward_f64(inp, v, h):
    """fp64 accumulation of the same sum: the 'truth' for protocol P2."""
    inp, v, h = _f32(inp), _f32(v), _f32(h)
    d, (B, C, H, W, K) = _dims(inp, v)
    out = np.empty((B, C, H, W), np.float64)
    _lib().oracle_sepconv_fwd_f64(_p(inp), _p(v), _p(h), _p(out), *d)
    return out


def sepconv_grad_vertical_reforder(g, inp, h):
    """kernel.cu:77-112, channel sum generalised from the literal 0,1,2 to C."""
    g, inp, h = _f32(g), _f32(inp), _f32(h)
    d, (B, C, H, W, K) = _dims(inp, h)
    gv = np.empty((B, K, H, W), np.float32)
    _lib().oracle_sepconv_gradv_reforder(_p(g), _p(inp), _p(h), _p(gv), *d)
    return gv


def sepconv_grad_horizontal_reforder(g, inp, v):
    """kernel.cu:115-150, channel sum generalised to C."""
    g, inp, v = _f32(g), _f32(inp), _f32(v)
    d, (B, C, H, W, K) = _dims(inp, v)
    gh = np.empty((B, K, H, W), np.float32)
    _lib().oracle_sepconv_gradh_reforder(_p(g), _p(inp), _p(v), _p(gh), *d)
    return gh


def sepconv_grad_taps_f64(g, inp, v, h):
    g, inp, v, h = _f32(g), _f32(inp), _f32(v), _f32(h)
    d, (B, C, H, W, K) = _dims(inp, v)
    gv = np.empty((B, K, H, W), np.float64)
    gh = np.empty((B, K, H, W), np.float64)
    _lib().oracle_sepconv_gradvh_f64(_p(g), _p(inp), _p(v), _p(h), _p(gv), _p(gh), *d)
    return gv, gh


def sepconv_grad_input_f64(g, v, h):
    """Adjoint of the forward (no reference implementation)."""
    g, v, h = _f32(g), _f32(v), _f32(h)
    B, C, H, W = g.shape
    K = v.shape[1]
    gi = np.empty((B, C, H + K - 1, W + K - 1), np.float64)
    _lib().oracle_sepconv_gradin_f64(_p(g), _p(v), _p(h), _p(gi), None,
                                     ctypes.c_int64(B), ctypes.c_int64(C), ctypes.c_int64(H),
                                     ctypes.c_int64(W), ctypes.c_int(K))
    return gi


def sepconv_fwd_bwd_fast(inp, v, h, g=None):
    """Factored fp32 CPU port, OpenMP over pixels: the cpu_baseline ('port')."""
    inp, v, h = _f32(inp), _f32(v), _f32(h)
    d, (B, C, H, W, K) = _dims(inp, v)
    out = np.empty((B, C, H, W), np.float32)
    if g is None:
        _lib().oracle_sepconv_fwd_bwd_fast(None, _p(inp), _p(v), _p(h), _p(out), None, None, *d)
        return out
    g = _f32(g)
    gv = np.empty((B, K, H, W), np.float32)
    gh = np.empty((B, K, H, W), np.float32)
    _lib().oracle_sepconv_fwd_bwd_fast(_p(g), _p(inp), _p(v), _p(h), _p(out), _p(gv), _p(gh), *d)
    return out, gv, gh


def sepconv_unfold_torch(inp, v, h):
    """The CPU baseline BASELINE.json names for sepconv: an unfold-based torch-CPU
    evaluation of kernel.cu:45-49 (row windows dotted with h, then with v,
    accumulated over fy).  Differentiable, so fwd+bwd is timed through autograd."""
    import torch

    K = v.shape[1]
    H, W = v.shape[2], v.shape[3]
    out = None
    for fy in range(K):
        rows = inp[:, :, fy:fy + H, :]                      # [B,C,H,W+K-1]
        win = rows.unfold(3, K, 1)                          # [B,C,H,W,K]
        r = (win * h.permute(0, 2, 3, 1).unsqueeze(1)).sum(-1)
        term = r * v[:, fy].unsqueeze(1)
        out = term if out is None else out + term
    return out


# --------------------------------------------------------------------------- interpolation tail
def _replicate_pad(x, pad=25):
    """nn.ReplicationPad2d(pad) (model_interp.py:46) == numpy edge padding of the last two axes."""
    return np.pad(_f32(x), ((0, 0), (0, 0), (pad, pad), (pad, pad)), mode="edge")


def interp_tail_reference(i1, i2, k1v, k1h, k2v, k2h):
    """The expression IFNet.forward ends with (sff_scripts_interp/model/model_interp.py:90-97),
    restated from the pinned pieces: replicate-pad both frames, the reference-order sepconv of
    each (kernel.cu:38-51), frame 2's result + frame 1's, then torch.mean(dim=1, keepdim=True)
    (fp32 sum over channels divided by C).  -> [B,1,H,W] float32."""
    y = sepconv_forward_reforder(_replicate_pad(i2), k2v, k2h) + sepconv_forward_reforder(_replicate_pad(i1), k1v, k1h)
    acc = y[:, 0].copy()
    for c in range(1, y.shape[1]):
        acc = acc + y[:, c]
    return (acc / np.float32(y.shape[1]))[:, None]


def interp_tail_f64(i1, i2, k1v, k1h, k2v, k2h):
    y = sepconv_forward_f64(_replicate_pad(i2), k2v, k2h) + sepconv_forward_f64(_replicate_pad(i1), k1v, k1h)
    return y.mean(axis=1, keepdims=True)


def interp_tail_grads_f64(g, i1, i2, k1v, k1h, k2v, k2h):
    """fp64 tap gradients of the tail given g [B,1,H,W]: the mean hands g/C to every channel, the
    add hands it to both sepconvs (kernel.cu:97-111, :134-149 with that upstream gradient).
    -> (g_k1v, g_k1h, g_k2v, g_k2h)."""
    C = np.asarray(i1).shape[1]
    gc = np.repeat(_f32(g) / np.float32(C), C, axis=1)
    g1v, g1h = sepconv_grad_taps_f64(gc, _replicate_pad(i1), k1v, k1h)
    g2v, g2h = sepconv_grad_taps_f64(gc, _replicate_pad(i2), k2v, k2h)
    return g1v, g1h, g2v, g2h


# --------------------------------------------------------------------------- warps
def warp_torch_restated(moving, flow):
    """numpy-fp32 restatement of SpatialTransformation.forward
    (sff_scripts_unfolding/utils/image_warp_torch.py:97-113, interpolate :32-95).

    moving [B,C,H,W] float32, flow [B,H,W,2] float32 (ch0 = x, ch1 = y) ->
    [B,C,H,W] float32.  Op order kept: x = (fx + j) + 1; x0 = floor(x); x1 = x0+1;
    clamp both to [0, W+1] (the 1-px zero-padded image); dx = float(x1c) - x;
    wa = dx*dy, wb = dx*(1-dy), wc = (1-dx)*dy, wd = (1-dx)*(1-dy);
    out = ((wa*Ia + wb*Ib) + wc*Ic) + wd*Id with Ia=(y0,x0) Ib=(y1,x0) Ic=(y0,x1) Id=(y1,x1).
    """
    moving = _f32(moving)
    flow = np.asarray(flow, dtype=np.float32)
    B, C, H, W = moving.shape
    one = np.float32(1.0)
    jj = np.arange(W, dtype=np.float32)[None, None, :]
    ii = np.arange(H, dtype=np.float32)[None, :, None]
    x = (flow[..., 0] + jj) + one
    y = (flow[..., 1] + ii) + one
    x0 = np.floor(x).astype(np.int64)
    y0 = np.floor(y).astype(np.int64)
    x1 = x0 + 1
    y1 = y0 + 1
    x0 = np.clip(x0, 0, W + 1); x1 = np.clip(x1, 0, W + 1)
    y0 = np.clip(y0, 0, H + 1); y1 = np.clip(y1, 0, H + 1)
    dx = x1.astype(np.float32) - x
    dy = y1.astype(np.float32) - y
    wa = dx * dy
    wb = dx * (one - dy)
    wc = (one - dx) * dy
    wd = (one - dx) * (one - dy)
    pad = np.zeros((B, C, H + 2, W + 2), np.float32)
    pad[:, :, 1:-1, 1:-1] = moving
    bi = np.arange(B)[:, None, None]
    out = np.empty((B, C, H, W), np.float32)
    for c in range(C):
        pc = pad[:, c]
        Ia = pc[bi, y0, x0]; Ib = pc[bi, y1, x0]; Ic = pc[bi, y0, x1]; Id = pc[bi, y1, x1]
        out[:, c] = ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id
    return out


def image_warp_restated(im, flow, mode="bilinear", return_float=False):
    """Restatement of numpy image_warp (simu_sff/image_warp.py:3-111).

    im ndim 2/3/4 = [[B],H,W,[C]], flow [[B],H,W,2].  x0 = clip(j + floor(fx), 0, W-1);
    x1 = clip(x0 + 1, 0, W-1) computed from the CLIPPED x0 (:84-88); weights from
    frac(flow) regardless of clipping (:72-82); 'nearest' samples (y0,x0), i.e.
    floor (:67-69); result truncated to uint8 (:110).  `return_float` also returns
    the value before the cast (used for bit-parity of the CUDA kernel).
    """
    im = np.asarray(im)
    flow = np.asarray(flow)
    nd = im.ndim
    if nd == 2:
        im4, fl4 = im[None, :, :, None], flow[None]
    elif nd == 3:
        im4, fl4 = im[None], flow[None]
    elif nd == 4:
        im4, fl4 = im, flow
    else:
        raise AttributeError("The dimension of im must be 2, 3 or 4")
    B, H, W, C = im4.shape
    fl_floor = np.floor(fl4)
    fi = fl_floor.astype(np.int32)
    jj = np.arange(W)[None, None, :]
    ii = np.arange(H)[None, :, None]
    x0 = np.clip(jj + fi[..., 0], 0, W - 1)
    y0 = np.clip(ii + fi[..., 1], 0, H - 1)
    bi = np.arange(B)[:, None, None]
    if mode == "nearest":
        val = im4[bi, y0, x0]
    elif mode == "bilinear":
        frac = fl4 - fl_floor
        xw = frac[..., 0][..., None]
        yw = frac[..., 1][..., None]
        x1 = np.clip(x0 + 1, 0, W - 1)
        y1 = np.clip(y0 + 1, 0, H - 1)
        wa = (1 - xw) * (1 - yw)
        wb = (1 - xw) * yw
        wc = xw * (1 - yw)
        wd = xw * yw
        val = wa * im4[bi, y0, x0] + wb * im4[bi, y1, x0] + wc * im4[bi, y0, x1] + wd * im4[bi, y1, x1]
    else:
        raise UnboundLocalError("mode must be 'nearest' or 'bilinear'")
    valf = val
    if nd == 2:
        valf = np.squeeze(valf)
    elif nd == 3:
        valf = np.squeeze(valf, axis=0)
    u8 = valf.astype(np.uint8)
    return (u8, valf) if return_float else u8


# --------------------------------------------------------------------------- SFF simulation (config 1)
def _sff_border_point(side, height, width, offset, crop_size, rng):
    x = rng.randint(1, (width if side in (1, 3) else height) - 1)
    while x < offset or x > crop_size - 50:
        x = rng.randint(1, width - 1)
    return {1: [0, x], 2: [x, width], 3: [height, x], 4: [x, 0]}[side]


def sff_get_two_points(height, width, offset, crop_size, rng):
    """simu_sff/simuSFF.py:42-94, same sequence of random draws."""
    k1 = rng.randint(1, 4)
    k2 = rng.randint(1, 4)
    while k1 == k2:
        k2 = rng.randint(1, 4)
    return (_sff_border_point(k1, height, width, offset, crop_size, rng),
            _sff_border_point(k2, height, width, offset, crop_size, rng))


def sff_degradation_restated(img, crop_size, rng, offset=50):
    """simu_sff/simuSFF.py:96-132 restated on numpy: draw a fold line, gen_flow
    (flow_synthesis.py:27-83, via the golden-pinned restatement in synth.gen_flow), numpy
    image_warp (image_warp_restated), mask multiply, uint8 cast; repeat until >= 100 zero pixels.
    -> (deformed uint8, flow float32 [H,W,2], mask float64)."""
    import math
    from sstem_restoration_b200 import synth            # numpy-only restatement, pinned by gen_flow_ref.npz
    while True:
        height = width = crop_size
        line_width = rng.randint(5, 20)
        fold_width = rng.randint(10, 80)
        dist = lambda a, c: math.sqrt((a[0] - c[0]) ** 2 + (a[1] - c[1]) ** 2)
        p1, p2 = sff_get_two_points(height, width, offset, crop_size, rng)
        while dist(p1, p2) < crop_size / 2:
            p1, p2 = sff_get_two_points(height, width, offset, crop_size, rng)
        dis_k = rng.uniform(0.00001, 0.1)
        k, b = synth.gen_line(p1, p2)
        flow, mask = synth.gen_flow(height, width, k, b, line_width, fold_width, dis_k)
        deformed = image_warp_restated(img, flow, mode="bilinear")
        deformed = (deformed * mask).astype(np.uint8)
        if len(np.where(deformed == 0)[0]) >= 100:
            return deformed, flow, mask


def sff_noise_restated(img, det_size, rng):
    """simu_sff/simuSFF.py:134-144 (works on a copy; the reference mutates its argument's box)."""
    img = np.array(img, copy=True)
    mask = np.ones_like(img)
    mask[img == 0] = 0
    ran = rng.uniform(0.4, 1.0)
    ran_w = rng.randint(50, 200)
    ran_h = rng.randint(50, 200)
    px = rng.randint(0, det_size - ran_h)
    py = rng.randint(0, det_size - ran_w)
    box = img[px:px + ran_h, py:py + ran_w]
    box = ran * (box - np.mean(img)) + np.mean(img)
    img[px:px + ran_h, py:py + ran_w] = box
    return np.multiply(img, mask)


def provider_degradation_restated(img, crop_size, offset, rng, line_width_max=50):
    """Provider.degradation of the training data providers restated on numpy
    (sff_scripts_unfolding/data/data_provider.py:180-245; fusion: line_width_max = 20).
    -> (deformed uint8 centre crop, flow2 float32 centre crop)."""
    from sstem_restoration_b200 import synth
    while True:
        height = width = crop_size
        line_width = rng.randint(5, line_width_max)
        fold_width = rng.randint(line_width + 1, 80)
        k1 = rng.randint(1, 4)
        k2 = rng.randint(1, 4)
        while k1 == k2:
            k2 = rng.randint(1, 4)
        pts = []
        for side in (k1, k2):
            x = rng.randint(1, (width if side in (1, 3) else height) - 1)
            pts.append({1: [0, x], 2: [x, width], 3: [height, x], 4: [x, 0]}[side])
        dis_k = rng.uniform(0.00001, 0.1)
        k, b = synth.gen_line(pts[0], pts[1])
        flow, flow2, mask = synth.gen_flow(height, width, k, b, line_width, fold_width, dis_k, two_flows=True)
        deformed = (image_warp_restated(img, flow, mode="bilinear") * mask).astype(np.uint8)
        deformed = deformed[offset:-offset, offset:-offset]
        flow2 = flow2[offset:-offset, offset:-offset]
        if len(np.where(deformed == 0)[0]) >= 100:
            return deformed, flow2


# --------------------------------------------------------------------------- stack pre/post-processing
def sections_to_input_restated(img1, img2, pad):
    """sff_scripts_interp/inference.py:69-83 on numpy (F.pad with zeros == np.pad constant)."""
    img1 = np.repeat(np.asarray(img1)[np.newaxis, :, :], 3, 0)
    img2 = np.repeat(np.asarray(img2)[np.newaxis, :, :], 3, 0)
    inputs = np.concatenate([img1, img2], axis=0)[np.newaxis]
    inputs = inputs.astype(np.float32) / 255.0
    return np.pad(inputs, ((0, 0), (0, 0), (pad, pad), (pad, pad)))


def prediction_to_uint8_restated(pred, pad):
    """inference.py:84-88: negative pad = crop, squeeze, * 255, astype(uint8)."""
    pred = np.asarray(pred, dtype=np.float32)
    if pad:
        pred = pred[..., pad:-pad, pad:-pad]
    return (np.squeeze(pred) * 255).astype(np.uint8)


def warp_stitch_restated(warped, img_interp):
    """sff_scripts_fusion/inference.py:163-171 restated in numpy (test oracle):

        warped_sff = (np.squeeze(warped) * 255).astype(np.uint8)             # [C,H,W]
        warped_sff = np.asarray(Image.fromarray(warped_sff.transpose(1,2,0)).convert('L'))
        mask = np.ones_like(warped_sff, np.float32); mask[warped_sff < 2] = 0
        stitch = (img_interp * (1 - mask) + warped_sff * mask).astype(np.uint8)

    ``warped``: float32 [C,H,W] (C = 1 or 3), ``img_interp``: uint8 [H,W].  Pillow's RGB -> 'L' is the fixed-point ITU-R
    601 luma ``(19595 R + 38470 G + 7471 B + 0x8000) >> 16`` (Pillow src/libImaging/Convert.c, macro L24); a CPU test pins
    this restatement against Pillow itself.  Returns (warped_gray uint8 [H,W], stitch uint8 [H,W])."""
    w8 = (np.asarray(warped, np.float32) * np.float32(255)).astype(np.uint8)
    if w8.shape[0] == 3:
        r, g, b = (w8[i].astype(np.uint32) for i in range(3))
        gray = ((19595 * r + 38470 * g + 7471 * b + 0x8000) >> 16).astype(np.uint8)
    else:
        gray = w8[0]
    mask = np.ones_like(gray, dtype=np.float32)
    mask[gray < 2] = 0
    stitch = (img_interp * (1 - mask) + gray * mask).astype(np.uint8)
    return gray, stitch


def warp_torch_backward_restated(moving, flow, grad_out):
    """Gradients of SpatialTransformation.forward (image_warp_torch.py:32-95) w.r.t. the moving image and the flow, as
    autograd derives them through the reference's ATen ops (test oracle; float64 accumulation):

      out = wa*Ia + wb*Ib + wc*Ic + wd*Id,  wa = dx*dy, wb = dx*(1-dy), wc = (1-dx)*dy, wd = (1-dx)*(1-dy),
      dx = x1_clamped - x, dy = y1_clamped - y  (floor / clamp / the gather indices carry no gradient), so
      d out / d fx = -(d out / d dx) = dy*(Ic - Ia) + (1-dy)*(Id - Ib),   d out / d fy = dx*(Ib - Ia) + (1-dx)*(Id - Ic),
      d out / d I(y,x) = the weight of every tap that reads it (taps in the 1-px zero border give nothing).

    moving [B,C,H,W], flow [B,H,W,2], grad_out [B,C,H,W] -> (grad_moving [B,C,H,W], grad_flow [B,H,W,2]) float32.
    Pinned against the reference's own autograd by tests/golden/warp_torch_grad_ref.npz."""
    moving = _f32(moving)
    flow = np.asarray(flow, dtype=np.float32)
    g = np.asarray(grad_out, dtype=np.float64)
    B, C, H, W = moving.shape
    one = np.float32(1.0)
    jj = np.arange(W, dtype=np.float32)[None, None, :]
    ii = np.arange(H, dtype=np.float32)[None, :, None]
    x = (flow[..., 0] + jj) + one
    y = (flow[..., 1] + ii) + one
    x0 = np.clip(np.floor(x).astype(np.int64), 0, W + 1); x1 = np.clip(np.floor(x).astype(np.int64) + 1, 0, W + 1)
    y0 = np.clip(np.floor(y).astype(np.int64), 0, H + 1); y1 = np.clip(np.floor(y).astype(np.int64) + 1, 0, H + 1)
    dx = (x1.astype(np.float32) - x).astype(np.float64)
    dy = (y1.astype(np.float32) - y).astype(np.float64)
    wa, wb, wc, wd = dx * dy, dx * (1 - dy), (1 - dx) * dy, (1 - dx) * (1 - dy)
    pad = np.zeros((B, C, H + 2, W + 2), np.float64)
    pad[:, :, 1:-1, 1:-1] = moving
    gpad = np.zeros_like(pad)
    gflow = np.zeros((B, H, W, 2), np.float64)
    bi = np.arange(B)[:, None, None] + np.zeros((B, H, W), np.int64)
    for c in range(C):
        pc = pad[:, c]
        Ia, Ib, Ic, Id = pc[bi, y0, x0], pc[bi, y1, x0], pc[bi, y0, x1], pc[bi, y1, x1]
        gc = g[:, c]
        gflow[..., 0] += gc * (dy * (Ic - Ia) + (1 - dy) * (Id - Ib))
        gflow[..., 1] += gc * (dx * (Ib - Ia) + (1 - dx) * (Id - Ic))
        for w, yy, xx in ((wa, y0, x0), (wb, y1, x0), (wc, y0, x1), (wd, y1, x1)):
            np.add.at(gpad[:, c], (bi, yy, xx), w * gc)
    return gpad[:, :, 1:-1, 1:-1].astype(np.float32), gflow.astype(np.float32)


def upsample2x_align_corners_restated(x, dtype=np.float64):
    """``nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)`` (sff_scripts_interp/model/model_interp.py:18)
    the way ATen's upsample_bilinear2d computes it for float tensors: scale = (in - 1) / (out - 1) in float32, source index
    = scale * dst (float32) cut to int, neighbour = + 1 unless that is past the last row / column, weights from the
    fraction (float32); value = h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d).  Indices and weights in float32 as ATen;
    the blend in ``dtype`` (float64: ground truth for the TF32 bound; float32: comparable with the golden ``up`` to an ulp).
    Pinned by tests/golden/tap_producer_ref.npz (made from the reference's own model on CPU)."""
    x = np.asarray(x)
    B, C, h, w = x.shape
    H, W = 2 * h, 2 * w

    def axis(n_in, n_out):
        scale = np.float32(n_in - 1) / np.float32(n_out - 1) if n_out > 1 else np.float32(0)
        r = (scale * np.arange(n_out, dtype=np.float32)).astype(np.float32)
        i0 = r.astype(np.int64)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (r - i0.astype(np.float32)).astype(np.float32)
        l0 = (np.float32(1) - l1).astype(np.float32)
        return i0, i1, l0.astype(dtype), l1.astype(dtype)

    y0, y1, hy0, hy1 = axis(h, H)
    x0, x1, wx0, wx1 = axis(w, W)
    xd = x.astype(dtype)
    top = wx0 * xd[:, :, y0][:, :, :, x0] + wx1 * xd[:, :, y0][:, :, :, x1]
    bot = wx0 * xd[:, :, y1][:, :, :, x0] + wx1 * xd[:, :, y1][:, :, :, x1]
    return hy0[:, None] * top + hy1[:, None] * bot


def tap_conv3x3_restated(x, weight, bias=None, upsample=True, dtype=np.float64):
    """The tail of ``IFNet._kernel_module`` (model_interp.py:130-137): [upsample x2 ->] ``Conv2d(cin, cout, 3, 1, 1)``
    (cross-correlation, zero padding 1), accumulated in ``dtype``.  x [B,cin,h,w], weight [cout,cin,3,3], bias [cout]."""
    up = upsample2x_align_corners_restated(x, dtype) if upsample else np.asarray(x).astype(dtype)
    B, C, H, W = up.shape
    wt = np.asarray(weight).astype(dtype)
    pad = np.zeros((B, C, H + 2, W + 2), dtype)
    pad[:, :, 1:-1, 1:-1] = up
    out = np.zeros((B, wt.shape[0], H, W), dtype)
    for dy in range(3):
        for dx in range(3):
            out += np.einsum("nc,bchw->bnhw", wt[:, :, dy, dx], pad[:, :, dy:dy + H, dx:dx + W], optimize=True)
    if bias is not None:
        out += np.asarray(bias).astype(dtype)[None, :, None, None]
    return out

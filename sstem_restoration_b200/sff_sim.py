"""GPU-resident SFF (support-film fold) simulation: the reference's ``simu_sff`` functions over
the sm_100a C ABI.

Mirrors simu_sff/simuSFF.py of the reference, function for function:

  * ``get_two_points`` (:42-94), ``cal_distance`` (:39-40), ``gen_line`` (flow_synthesis.py:13-19) --
    host-side scalar logic, restated with the same sequence of ``random`` draws so that a seeded
    run picks the same fold line as the reference;
  * ``degradation(img, crop_size, offset=50)`` (:96-132) -> ``(deformed, flow, mask)``:
    ``gen_flow`` + ``image_warp`` + mask multiply + zero count run as ONE kernel
    (``sstem_sff_degrade``); the accept loop (``count < 100`` -> draw again) stays on the host and
    reads back 8 bytes per attempt;
  * ``noise(img, det_size)`` (:134-144): ``sstem_sff_contrast`` on the box only;
  * ``SimuSFF`` without the PNG I/O: :func:`simu_sff` returns the degraded patch and the flow.

Images are CUDA uint8 tensors ``[H, W]`` (or numpy arrays, which are uploaded and the results
downloaded -- still the CUDA path).  Results are bit-equal to the reference's numpy run for the
same ``random`` state.  There is no CPU fallback.
"""
from __future__ import annotations

import math
import random as _random

import numpy as np
import torch

from . import _lib

_MINA = 0.000000001                                      # flow_synthesis.py:10


def gen_line(p1, p2):
    """flow_synthesis.py:13-19."""
    denominator = p2[1] - p1[1]
    if denominator == 0:
        denominator = _MINA
    k = (p2[0] - p1[0]) / denominator
    return k, p1[0] - (k * p1[1])


def cal_distance(p1, p2):
    return math.sqrt((p1[0] - p2[0]) ** 2 + (p1[1] - p2[1]) ** 2)


def _border_point(side, height, width, offset, crop_size, rng):
    # simuSFF.py:52-93: the first draw spans the side's own length, every redraw `width - 1`
    x = rng.randint(1, (width if side in (1, 3) else height) - 1)
    while x < offset or x > crop_size - 50:
        x = rng.randint(1, width - 1)
    return {1: [0, x], 2: [x, width], 3: [height, x], 4: [x, 0]}[side]


def get_two_points(height, width, offset, crop_size, rng=_random):
    """simuSFF.py:42-94: two end points on two different borders (1 top, 2 right, 3 bottom, 4 left)."""
    k1 = rng.randint(1, 4)
    k2 = rng.randint(1, 4)
    while k1 == k2:
        k2 = rng.randint(1, 4)
    p1 = _border_point(k1, height, width, offset, crop_size, rng)
    p2 = _border_point(k2, height, width, offset, crop_size, rng)
    return p1, p2


def fold_line_params(k, b, line_width, fold_width, dis_k):
    """The 8 scalars ``sstem_sff_degrade`` takes per image, derived with Python's ``math`` exactly
    as flow_synthesis.py:31,64-71 derives them."""
    k_T = 1 / _MINA if k == 0 else 1 / k
    angle = math.atan(k_T)
    return [float(k), float(b), math.sqrt(k ** 2 + 1), float(line_width), float(fold_width), float(dis_k),
            math.sin(angle), math.cos(angle)]


def _as_cuda_u8(img):
    if not torch.cuda.is_available():
        raise _lib.SstemError("sstem_restoration_b200 sff_sim: no CUDA device; there is no CPU fallback")
    host = not (isinstance(img, torch.Tensor) and img.is_cuda)
    t = torch.as_tensor(img)
    if t.dtype != torch.uint8:
        raise TypeError("sff_sim: uint8 grayscale image required (skimage.io.imread of a section, simuSFF.py:16)")
    if t.dim() not in (2, 3):
        raise ValueError("sff_sim: image must be [H,W] or [B,H,W]")
    if host:
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    return t.contiguous(), host


def gen_flow_warp(img, params, want_flow=True, want_mask=True, want_flow2=False, stats_border=0):
    """Kernel-level entry: ``img`` CUDA uint8 [B,H,W], ``params`` list of B 8-tuples
    (:func:`fold_line_params`).  -> (deformed [B,H,W] u8, flow [B,H,W,2] f32 | None,
    mask [B,H,W] u8 | None, stats [B,2] int64 CUDA: zero count, pixel sum) and, with
    ``want_flow2``, the data providers' second flow as a fifth element."""
    B, H, W = img.shape
    dev = img.device
    p = torch.tensor(params, dtype=torch.float64).reshape(B, 8).to(dev)
    out = torch.empty_like(img)
    flow = torch.empty((B, H, W, 2), dtype=torch.float32, device=dev) if want_flow else None
    flow2 = torch.empty((B, H, W, 2), dtype=torch.float32, device=dev) if want_flow2 else None
    mask = torch.empty_like(img) if want_mask else None
    stats = torch.empty((B, 2), dtype=torch.int64, device=dev)
    code = _lib.load().sstem_sff_degrade(
        img.data_ptr(), p.data_ptr(), out.data_ptr(), flow.data_ptr() if want_flow else None,
        flow2.data_ptr() if want_flow2 else None, mask.data_ptr() if want_mask else None, stats.data_ptr(),
        B, H, W, stats_border, torch.cuda.current_stream(dev).cuda_stream)
    if code:
        _lib.check(code, "sstem_sff_degrade")
    return (out, flow, mask, stats, flow2) if want_flow2 else (out, flow, mask, stats)


def gen_flow(height, width, k, b, line_width=5, fold_width=10, dis_k=0.1, two_flows=False):
    """Drop-in for ``gen_flow`` (simu_sff/flow_synthesis.py:27-83; with ``two_flows`` the data providers'
    variant, sff_scripts_unfolding/utils/flow_synthesis.py:27-61, which also returns ``flow2``), evaluated by
    the kernel: -> (flow float32 [H,W,2], [flow2,] mask float64 [H,W]) as numpy arrays, bit-equal."""
    if not torch.cuda.is_available():
        raise _lib.SstemError("sstem_restoration_b200 sff_sim: no CUDA device; there is no CPU fallback")
    img = torch.zeros((1, height, width), dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
    res = gen_flow_warp(img, [fold_line_params(k, b, line_width, fold_width, dis_k)], want_flow2=two_flows)
    flow, mask = res[1][0].cpu().numpy(), res[2][0].cpu().numpy().astype(np.float64)
    return (flow, res[4][0].cpu().numpy(), mask) if two_flows else (flow, mask)


def _draw_provider_line(crop_size, rng, line_width_max):
    """One attempt's draws of Provider.degradation (data_provider.py:185-222), in the reference's order."""
    height = width = crop_size
    line_width = rng.randint(5, line_width_max)
    fold_width = rng.randint(line_width + 1, 80)
    k1 = rng.randint(1, 4)
    k2 = rng.randint(1, 4)
    while k1 == k2:
        k2 = rng.randint(1, 4)
    pts = []
    for side in (k1, k2):                                 # data_provider.py:195-219: one draw per point, no offset rule
        x = rng.randint(1, (width if side in (1, 3) else height) - 1)
        pts.append({1: [0, x], 2: [x, width], 3: [height, x], 4: [x, 0]}[side])
    dis_k = rng.uniform(0.00001, 0.1)
    k, b = gen_line(pts[0], pts[1])
    return fold_line_params(k, b, line_width, fold_width, dis_k)


def _draw_noise_box(det_size, rng):
    """The draws of `noise` (simuSFF.py:137-141 / data_provider.py:249-254) -> the 8 scalars of sstem_sff_contrast."""
    ran = rng.uniform(0.4, 1.0)
    ran_w = rng.randint(50, 200)
    ran_h = rng.randint(50, 200)
    px = rng.randint(0, det_size - ran_h)
    py = rng.randint(0, det_size - ran_w)
    return [ran, px, py, ran_h, ran_w, 0, 0, 0]


def _launch_degrade_batch(imgs, params, offset):
    """imgs [n,crop,crop] CUDA u8 -> (deformed centre crops [n,det,det] u8, flow2 crops, stats [n,2], zero counts list)."""
    out, _, _, stats, flow2 = gen_flow_warp(imgs, params, want_flow=False, want_mask=False, want_flow2=True, stats_border=offset)
    sl = slice(offset, -offset) if offset else slice(None)
    return out[:, sl, sl].contiguous(), flow2[:, sl, sl].contiguous(), stats, stats[:, 0].tolist()


def _launch_contrast_batch(imgs, stats, boxes):
    """In place on imgs [n,det,det] CUDA u8."""
    n, H, W = imgs.shape
    p = torch.tensor(boxes, dtype=torch.float64).reshape(n, 8).to(imgs.device)
    code = _lib.load().sstem_sff_contrast(imgs.data_ptr(), stats.contiguous().data_ptr(), p.data_ptr(), n, H, W,
                                          max(int(b[3]) for b in boxes), max(int(b[4]) for b in boxes),
                                          torch.cuda.current_stream(imgs.device).cuda_stream)
    if code:
        _lib.check(code, "sstem_sff_contrast")
    return imgs


def provider_batch(imgs, crop_size, offset, rng=_random, line_width_max=50, with_noise=True):
    """A whole training batch through ``Provider.degradation`` (+ ``Provider.noise``) of the data providers
    (data_provider.py:180-259) with ONE degrade launch, one contrast launch and one 8*B-byte read in the common case,
    yet the same ``random`` draws -- hence bit-identical outputs -- as B sequential per-sample calls: the batch is
    drawn optimistically (every sample accepted at its first attempt); if sample i is rejected, the samples before it
    are kept, the generator is rewound to sample i's state with its failed attempt consumed, and the rest is redone.
    ``imgs``: CUDA uint8 [B, crop_size, crop_size].  -> (sff uint8 [B,det,det], flow2 float32 [B,det,det,2]),
    det = crop_size - 2*offset."""
    t, host = _as_cuda_u8(imgs)
    if t.dim() != 3:
        raise ValueError("provider_batch: images as [B,H,W]")
    B = t.shape[0]
    det = crop_size - 2 * offset
    sff, flow2 = [], []
    start = 0
    while start < B:
        states, lines, boxes = [], [], []
        for _ in range(start, B):
            states.append(rng.getstate())
            lines.append(_draw_provider_line(crop_size, rng, line_width_max))
            if with_noise:
                boxes.append(_draw_noise_box(det, rng))
        out, fl2, stats, zeros = _launch_degrade_batch(t[start:], lines, offset)
        bad = next((i for i, z in enumerate(zeros) if z < 100), None)     # data_provider.py:236-241
        n_ok = len(zeros) if bad is None else bad
        if n_ok:
            good = out[:n_ok]
            if with_noise:
                good = _launch_contrast_batch(good.contiguous(), stats[:n_ok], boxes[:n_ok])
            sff.append(good)
            flow2.append(fl2[:n_ok])
        if bad is not None:
            rng.setstate(states[bad])
            _draw_provider_line(crop_size, rng, line_width_max)             # the failed attempt's draws stay consumed
        start += n_ok
    sff, flow2 = torch.cat(sff, 0), torch.cat(flow2, 0)
    if host:
        return sff.cpu().numpy(), flow2.cpu().numpy()
    return sff, flow2


def provider_degradation(img, crop_size, offset, rng=_random, line_width_max=50):
    """``Provider.degradation`` of the training data providers on the GPU
    (sff_scripts_unfolding/data/data_provider.py:180-245; the fusion provider is the same with
    ``line_width_max=20``, sff_scripts_fusion/data/data_provider.py:188): fold line through two random
    border points, both flows, warp, mask, centre crop by ``offset``, accepted once the crop holds >= 100
    zero pixels.  -> (deformed uint8 [crop-2*offset]^2, flow2 float32 [.., .., 2])."""
    t, host = _as_cuda_u8(img)
    if t.dim() != 2:
        raise ValueError("provider_degradation: one [H,W] image")
    t = t[None]
    while True:
        out, _, _, stats, flow2 = gen_flow_warp(t, [_draw_provider_line(crop_size, rng, line_width_max)],
                                                want_flow=False, want_mask=False, want_flow2=True, stats_border=offset)
        if int(stats[0, 0].item()) >= 100:
            break
    sl = slice(offset, -offset) if offset else slice(None)
    deformed, flow2 = out[0][sl, sl], flow2[0][sl, sl]
    if host:
        return deformed.cpu().numpy(), flow2.cpu().numpy()
    return deformed.contiguous(), flow2.contiguous()


def degradation(img, crop_size, offset=50, rng=_random, return_stats=False):
    """simuSFF.py:96-132 on the GPU.  ``img``: uint8 [crop_size, crop_size].
    -> (deformed uint8, flow float32 [H,W,2], mask) -- mask as float64 0/1 like the reference's
    when the input was a numpy array, uint8 0/1 on the device otherwise."""
    t, host = _as_cuda_u8(img)
    if t.dim() != 2:
        raise ValueError("degradation: one [H,W] image (the reference's patch)")
    t = t[None]
    while True:
        height = width = crop_size
        line_width = rng.randint(5, 20)
        fold_width = rng.randint(10, 80)
        p1, p2 = get_two_points(height, width, offset, crop_size, rng)
        while cal_distance(p1, p2) < crop_size / 2:
            p1, p2 = get_two_points(height, width, offset, crop_size, rng)
        dis_k = rng.uniform(0.00001, 0.1)
        k, b = gen_line(p1, p2)
        out, flow, mask, stats = gen_flow_warp(t, [fold_line_params(k, b, line_width, fold_width, dis_k)])
        if int(stats[0, 0].item()) >= 100:                # simuSFF.py:125-130 (one 8-byte read per attempt)
            break
    if host:
        res = (out[0].cpu().numpy(), flow[0].cpu().numpy(), mask[0].cpu().numpy().astype(np.float64))
    else:
        res = (out[0], flow[0], mask[0])
    return res + (stats,) if return_stats else res


def noise(img, det_size, rng=_random, stats=None):
    """simuSFF.py:134-144 on the GPU (regional contrast inside a random box; zero pixels stay zero).
    ``stats``: the [1,2] int64 tensor :func:`degradation` returned for this image (saves the
    reduction for np.mean); recomputed when absent.  Returns a new image, like the reference."""
    t, host = _as_cuda_u8(img)
    if t.dim() != 2:
        raise ValueError("noise: one [H,W] image")
    t = t.clone()[None]
    box = _draw_noise_box(det_size, rng)
    if stats is None:
        stats = torch.stack([(t == 0).sum(), t.sum(dtype=torch.int64)]).reshape(1, 2)
    stats = stats.to(device=t.device, dtype=torch.int64).contiguous()
    _launch_contrast_batch(t, stats, [box])
    return t[0].cpu().numpy() if host else t[0]


def simu_sff(clean_img, patch_size, rng=_random):
    """``SimuSFF`` (simuSFF.py:14-30) without the file I/O: random crop (when the image is larger
    than ``patch_size``), degradation, regional contrast.  -> (sff_patch uint8, flow, mask)."""
    t, host = _as_cuda_u8(clean_img)
    h, w = t.shape
    if patch_size < h and patch_size < w:
        i = rng.randint(0, h - patch_size)
        j = rng.randint(0, w - patch_size)
        patch, size = t[i:i + patch_size, j:j + patch_size].contiguous(), patch_size
    else:
        patch, size = t, h
    deformed, flow, mask, stats = degradation(patch, size, rng=rng, return_stats=True)
    out = noise(deformed, size, rng=rng, stats=stats)
    if host:
        return out.cpu().numpy(), flow.cpu().numpy(), mask.cpu().numpy().astype(np.float64)
    return out, flow, mask

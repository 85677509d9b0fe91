"""Synthetic inputs for tests and bench.py (host-side numpy; no kernels here).

  * EM-like sections: band-limited noise thresholded into
    thin dark membrane-like ridges on a mid-grey field, quantised to uint8.
  * ``gen_line`` / ``gen_flow``: the fold-line displacement field of the SFF
    simulator, restated from simu_sff/flow_synthesis.py:13-83 (that module cannot
    be imported as is -- it pulls in matplotlib at :6).
  * unit-scale taps (softmax-normalised) and the ReplicationPad2d(25) +
    gray->x3 packing the callers do (sff_scripts_interp/inference.py:71-77,
    model_interp.py:46,90-91).
"""
from __future__ import annotations

import math
import random

import numpy as np


def _gauss_blur_fft(img: np.ndarray, sigma: float) -> np.ndarray:
    h, w = img.shape
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.rfftfreq(w)[None, :]
    g = np.exp(-2.0 * (math.pi * sigma) ** 2 * (fx * fx + fy * fy))
    return np.fft.irfft2(np.fft.rfft2(img) * g, s=img.shape)


def em_section(height: int, width: int, index: int = 0) -> np.ndarray:
    """Deterministic uint8 [H,W] EM-like section (rng seed 1234 + index)."""
    rng = np.random.default_rng(1234 + index)
    field = _gauss_blur_fft(rng.standard_normal((height, width)), 6.0)
    field /= field.std() + 1e-12
    ridges = np.exp(-(field / 0.12) ** 2)                  # thin lines where the field crosses zero
    blobs = _gauss_blur_fft(rng.standard_normal((height, width)), 14.0)
    blobs /= blobs.std() + 1e-12
    img = 150.0 - 95.0 * ridges + 18.0 * blobs + 9.0 * rng.standard_normal((height, width))
    return np.clip(img, 0, 255).astype(np.uint8)


def section_to_input(sec_u8: np.ndarray, pad: int = 25) -> np.ndarray:
    """uint8 [H,W] -> float32 [3, H+2*pad, W+2*pad]: /255, gray->x3, replicate-pad."""
    x = sec_u8.astype(np.float32) / np.float32(255.0)
    x = np.pad(x, pad, mode="edge")
    return np.repeat(x[None], 3, axis=0)


def unit_taps(batch: int, k: int, height: int, width: int, seed: int = 4321) -> np.ndarray:
    """softmax(randn) taps [B,K,H,W] float32: |out| stays O(1) (parity protocol P1)."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((batch, k, height, width)).astype(np.float32)
    z -= z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


# ---- fold-line flow (simu_sff/flow_synthesis.py) ----------------------------------
_MINA = 0.000000001


def gen_line(p1, p2):
    """flow_synthesis.py:13-19: slope / intercept through two (row, col) points."""
    den = p2[1] - p1[1]
    if den == 0:
        den = _MINA
    k = (p2[0] - p1[0]) / den
    return k, p1[0] - k * p1[1]


def gen_flow(height, width, k, b, line_width=5, fold_width=10, dis_k=0.1, two_flows=False):
    """flow_synthesis.py:27-83 -> (flow float32 [H,W,2], mask float64 [H,W]); with ``two_flows`` the data
    providers' variant (sff_scripts_unfolding/utils/flow_synthesis.py:27-61) -> (flow, flow2, mask).

    Signed distance to the fold line; inside the line (|d| <= line_width) the
    displacement equals the distance, outside it decays linearly from
    fold_width - line_width with slope dis_k, floored at 0; direction normal to
    the line.
    """
    xs = np.arange(width)[None, :].repeat(height, 0).reshape(-1)
    ys = np.arange(height)[:, None].repeat(width, 1).reshape(-1)
    d = ((k * xs - ys + b) / math.sqrt(k ** 2 + 1)).reshape(height, width)
    sign = np.zeros_like(d)
    sign[d > 0] = 1
    sign[d < 0] = -1
    ad = np.abs(d)
    mask = np.zeros_like(d)
    mask[ad > line_width] = 1
    outside = np.ones_like(d)
    outside[ad < line_width] = 0
    slope = -dis_k
    off = (fold_width - line_width) - slope * line_width
    mag = slope * ad + off
    mag[mag < 0] = 0
    if two_flows:                                          # unfolding flow_synthesis.py:48-49,58,62
        beyond = np.ones_like(d)
        beyond[ad < fold_width] = 0
        d2 = (mag * beyond + ad * (1 - beyond)) * (-sign)
    mag = mag * outside + ad * (1 - outside)
    d = mag * sign
    k_t = 1 / _MINA if k == 0 else 1 / k
    ang = math.atan(k_t)
    s, c = math.sin(ang), math.cos(ang)

    def project(dd):
        fl = np.zeros((height, width, 2), dtype=np.float32)
        if k > 0:
            fl[:, :, 0] = dd * c
            fl[:, :, 1] = -(dd * s)
        else:
            fl[:, :, 0] = -(dd * c)
            fl[:, :, 1] = dd * s
        return fl

    if two_flows:
        return project(d), project(d2), mask
    return project(d), mask


def random_fold_flow(height, width, seed=555):
    """Fold flow with the SFF simulator's parameter ranges (simu_sff/simuSFF.py:96-112),
    `random.seed(555)` being the repo's random_seed (config ms_l1loss_decay.yaml:33)."""
    rnd = random.Random(seed)
    line_width = rnd.randint(5, 20)
    fold_width = rnd.randint(10, 80)
    dis_k = rnd.uniform(0.00001, 0.1)
    # a long chord between two different borders
    p1 = [0, rnd.randint(width // 4, 3 * width // 4)]
    p2 = [height, rnd.randint(width // 4, 3 * width // 4)]
    k, b = gen_line(p1, p2)
    return gen_flow(height, width, k, b, line_width, fold_width, dis_k)


def noise_flow(height, width, sigma=5.0, seed=7):
    """N(0, sigma px) i.i.d. flow: the cache-hostile warp case."""
    rng = np.random.default_rng(seed)
    return (sigma * rng.standard_normal((height, width, 2))).astype(np.float32)

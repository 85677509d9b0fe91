"""Seeded input definitions shared by the golden-vector generator and the tests.
Inputs are rebuilt from seeds, so the committed fixtures hold reference OUTPUTS only."""
import numpy as np


def _rng(seed):
    return np.random.default_rng(seed)


def _fold_like_flow(h, w, seed, amp):
    r = _rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    d = (0.37 * xx - yy + 0.4 * h) / np.float32(np.sqrt(0.37 ** 2 + 1))
    mag = np.clip(amp - 0.05 * np.abs(d), 0, None) * np.sign(d)
    f = np.stack([mag * 0.94, -mag * 0.35], -1).astype(np.float32)
    return f + (0.3 * r.standard_normal((h, w, 2))).astype(np.float32)


def image_warp_cases():
    """name -> (im, flow, mode) for numpy image_warp."""
    c = {}
    r = _rng(11)
    im = r.integers(0, 256, (48, 64), dtype=np.uint8)
    c["u8_2d_bilinear_noise"] = (im, (3.0 * r.standard_normal((48, 64, 2))).astype(np.float32), "bilinear")
    c["u8_2d_nearest_noise"] = (im, (3.0 * r.standard_normal((48, 64, 2))).astype(np.float32), "nearest")
    c["u8_2d_bilinear_fold"] = (im, _fold_like_flow(48, 64, 12, 20.0), "bilinear")
    # far outside the image on every side: exercises the clip and the x1-from-clipped-x0 quirk
    c["u8_2d_bilinear_far"] = (im, (40.0 * r.standard_normal((48, 64, 2))).astype(np.float32), "bilinear")
    im3 = r.integers(0, 256, (33, 47, 3), dtype=np.uint8)
    c["u8_3d_bilinear"] = (im3, (4.0 * r.standard_normal((33, 47, 2))).astype(np.float32), "bilinear")
    c["u8_3d_nearest"] = (im3, (4.0 * r.standard_normal((33, 47, 2))).astype(np.float32), "nearest")
    im4 = r.integers(0, 256, (2, 21, 19, 2), dtype=np.uint8)
    c["u8_4d_bilinear"] = (im4, (2.5 * r.standard_normal((2, 21, 19, 2))).astype(np.float32), "bilinear")
    imf = (255.0 * r.random((40, 40))).astype(np.float32)
    c["f32_2d_bilinear"] = (imf, (3.0 * r.standard_normal((40, 40, 2))).astype(np.float32), "bilinear")
    # integer-valued and zero flows (weights exactly 0/1)
    c["u8_2d_bilinear_integer"] = (im, np.round(3.0 * r.standard_normal((48, 64, 2))).astype(np.float32), "bilinear")
    c["u8_2d_bilinear_zero"] = (im, np.zeros((48, 64, 2), np.float32), "bilinear")
    c["u8_1x1"] = (np.array([[77]], np.uint8), np.array([[[0.25, -0.75]]], np.float32), "bilinear")
    return c


def warp_torch_cases():
    """name -> (moving [B,C,H,W] f32, flow [B,H,W,2] f32) for SpatialTransformation."""
    c = {}
    r = _rng(21)
    c["b1c3_noise"] = (r.random((1, 3, 40, 56), dtype=np.float32), (3.0 * r.standard_normal((1, 40, 56, 2))).astype(np.float32))
    c["b2c1_noise"] = (r.random((2, 1, 31, 29), dtype=np.float32), (2.0 * r.standard_normal((2, 31, 29, 2))).astype(np.float32))
    c["b1c3_far"] = (r.random((1, 3, 24, 24), dtype=np.float32), (30.0 * r.standard_normal((1, 24, 24, 2))).astype(np.float32))
    c["b1c2_fold"] = (r.random((1, 2, 48, 64), dtype=np.float32), _fold_like_flow(48, 64, 22, 25.0)[None])
    c["b1c3_integer"] = (r.random((1, 3, 20, 20), dtype=np.float32), np.round(2.0 * r.standard_normal((1, 20, 20, 2))).astype(np.float32))
    c["b1c1_zero"] = (r.random((1, 1, 17, 23), dtype=np.float32), np.zeros((1, 17, 23, 2), np.float32))
    c["b3c3_wide"] = (r.random((3, 3, 9, 130), dtype=np.float32), (1.5 * r.standard_normal((3, 9, 130, 2))).astype(np.float32))
    # (a single-pixel image is not a case: the reference's torch.squeeze at image_warp_torch.py:93
    #  collapses the pixel axis and raises IndexError there)
    c["b1c1_2x1"] = (np.array([[[[0.5], [0.25]]]], np.float32), np.array([[[[0.25, -0.5]], [[-0.5, 0.75]]]], np.float32))
    return c


def gen_flow_cases():
    """name -> (h, w, p1, p2, line_width, fold_width, dis_k)."""
    return {
        "pos_slope": (64, 80, [0, 20], [64, 60], 5, 30, 0.05),
        "neg_slope": (64, 80, [64, 10], [0, 70], 8, 40, 0.01),
        "vertical_line": (48, 48, [0, 24], [48, 24], 6, 20, 0.1),
        "horizontal_line": (48, 48, [24, 0], [24, 48], 10, 15, 0.001),
    }


def simu_sff_cases():
    """name -> (patch size, section index of synth.em_section, random.seed) for simuSFF.degradation + noise."""
    return {"p256_seed555": (256, 3, 555), "p256_seed1": (256, 4, 1), "p300_seed77": (300, 5, 77)}


def provider_degradation_cases():
    """name -> (crop size, offset, section index, random.seed, provider) for Provider.degradation + noise
    of the training data providers (det_size = crop - 2*offset = 256 as in data_provider.py:94-95)."""
    return {"unfolding_c320_seed3": (320, 32, 6, 3, "unfolding"), "fusion_c288_seed11": (288, 16, 7, 11, "fusion")}


def sepconv_cases():
    """name -> dict(B, C, H, W, seed, scale): seeded sepconv inputs (K = 51)."""
    return {
        "b1c3_16x16_unit": dict(B=1, C=3, H=16, W=16, seed=101, kind="unit"),
        "b2c3_9x13_randn": dict(B=2, C=3, H=9, W=13, seed=102, kind="randn"),
        "b2c3_1x1_gradcheck": dict(B=2, C=3, H=1, W=1, seed=103, kind="randn"),  # model_interp.py:109-119 shapes
        "b1c1_20x24_unit": dict(B=1, C=1, H=20, W=24, seed=104, kind="unit"),
    }


def sepconv_inputs(B, C, H, W, seed, kind, K=51):
    r = _rng(seed)
    inp = r.random((B, C, H + K - 1, W + K - 1), dtype=np.float32)
    if kind == "unit":
        z = r.standard_normal((2, B, K, H, W)).astype(np.float32)
        e = np.exp(z - z.max(axis=2, keepdims=True))
        t = (e / e.sum(axis=2, keepdims=True)).astype(np.float32)
        v, h = t[0], t[1]
    else:
        v = r.standard_normal((B, K, H, W)).astype(np.float32)
        h = r.standard_normal((B, K, H, W)).astype(np.float32)
        inp = r.standard_normal((B, C, H + K - 1, W + K - 1)).astype(np.float32)
    g = r.standard_normal((B, C, H, W)).astype(np.float32)
    return inp, v, h, g


def kpn_taps_case():
    """BASELINE config 2: random-init KPN (torch.manual_seed) on two synthetic 256x256 sections; crop of the taps kept."""
    return dict(size=256, section_a=11, section_b=12, torch_seed=0, crop_y=40, crop_x=200, crop=32)


def kpn_frames(p):
    """[1,6,H,W] float32 network input exactly as sff_scripts_interp/inference.py:69-77 builds it (gray x3, /255)."""
    from sstem_restoration_b200 import synth
    a = synth.em_section(p["size"], p["size"], p["section_a"]).astype(np.float32) / 255.0
    b = synth.em_section(p["size"], p["size"], p["section_b"]).astype(np.float32) / 255.0
    return np.concatenate([np.repeat(a[None], 3, 0), np.repeat(b[None], 3, 0)], 0)[None].astype(np.float32)

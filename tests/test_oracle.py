"""CPU: the oracle against the reference's own outputs (golden fixtures) and against
independent restatements (fp64, torch autograd)."""
import os

import numpy as np
import pytest
import torch

import oracle
from tests.golden import cases


# ------------------------------------------------------------------ warps vs reference outputs
def test_image_warp_restatement_matches_reference_bitwise(golden_dir):
    ref = np.load(os.path.join(golden_dir, "warp_numpy_ref.npz"))
    for name, (im, flow, mode) in cases.image_warp_cases().items():
        got = oracle.image_warp_restated(im, flow, mode)
        assert got.dtype == np.uint8 and got.shape == ref[name].shape, name
        assert np.array_equal(got, ref[name]), name


def test_image_warp_bad_ndim_raises_like_reference():
    with pytest.raises(AttributeError):
        oracle.image_warp_restated(np.zeros((2, 2, 2, 2, 2), np.uint8), np.zeros((2, 2, 2), np.float32))


def test_warp_torch_restatement_matches_reference_bitwise(golden_dir):
    ref = np.load(os.path.join(golden_dir, "warp_torch_ref.npz"))
    for name, (moving, flow) in cases.warp_torch_cases().items():
        got = oracle.warp_torch_restated(moving, flow)
        assert got.shape == ref[name].shape, name
        assert np.array_equal(got.view(np.uint32), ref[name].view(np.uint32)), name


def test_gen_flow_restatement_matches_reference_bitwise(golden_dir):
    from sstem_restoration_b200 import synth
    ref = np.load(os.path.join(golden_dir, "gen_flow_ref.npz"))
    for name, (h, w, p1, p2, lw, fw, dk) in cases.gen_flow_cases().items():
        k, b = synth.gen_line(p1, p2)
        assert np.array_equal(np.array([k, b]), ref[name + "_kb"]), name
        flow, mask = synth.gen_flow(h, w, k, b, lw, fw, dk)
        assert flow.dtype == np.float32
        assert np.array_equal(flow.view(np.uint32), ref[name + "_flow"].view(np.uint32)), name
        assert np.array_equal(mask.astype(np.uint8), ref[name + "_mask"]), name


# ------------------------------------------------------------------ sepconv oracle
def _torch_sepconv_f64(inp, v, h):
    """Independent shift-and-add restatement in fp64 (torch, differentiable)."""
    K, H, W = v.shape[1], v.shape[2], v.shape[3]
    out = 0
    for fy in range(K):
        r = 0
        for fx in range(K):
            r = r + inp[:, :, fy:fy + H, fx:fx + W] * h[:, fx].unsqueeze(1)
        out = out + r * v[:, fy].unsqueeze(1)
    return out


@pytest.mark.parametrize("name", list(cases.sepconv_cases()))
def test_sepconv_oracles_agree(name):
    p = cases.sepconv_cases()[name]
    inp, v, h, g = cases.sepconv_inputs(**p)
    ref32 = oracle.sepconv_forward_reforder(inp, v, h)
    ref64 = oracle.sepconv_forward_f64(inp, v, h)
    t = _torch_sepconv_f64(*(torch.from_numpy(a).double() for a in (inp, v, h))).numpy()
    assert np.allclose(ref64, t, rtol=1e-12, atol=1e-12)
    scale = max(1.0, np.abs(ref64).max())
    assert np.abs(ref32 - ref64).max() <= 2e-5 * scale * (50 if p["kind"] == "randn" else 1)
    fast = oracle.sepconv_fwd_bwd_fast(inp, v, h)
    assert np.abs(fast - ref64).max() <= 2e-5 * scale * (50 if p["kind"] == "randn" else 1)


@pytest.mark.parametrize("name", list(cases.sepconv_cases()))
def test_sepconv_gradient_oracles_match_autograd(name):
    p = cases.sepconv_cases()[name]
    inp, v, h, g = cases.sepconv_inputs(**p)
    ti, tv, th = (torch.from_numpy(a).double().requires_grad_(True) for a in (inp, v, h))
    out = _torch_sepconv_f64(ti, tv, th)
    out.backward(torch.from_numpy(g).double())
    gv64, gh64 = oracle.sepconv_grad_taps_f64(g, inp, v, h)
    gi64 = oracle.sepconv_grad_input_f64(g, v, h)
    assert np.allclose(gv64, tv.grad.numpy(), rtol=1e-11, atol=1e-11)
    assert np.allclose(gh64, th.grad.numpy(), rtol=1e-11, atol=1e-11)
    assert np.allclose(gi64, ti.grad.numpy(), rtol=1e-11, atol=1e-11)
    gv32 = oracle.sepconv_grad_vertical_reforder(g, inp, h)
    gh32 = oracle.sepconv_grad_horizontal_reforder(g, inp, v)
    sv, sh = max(1.0, np.abs(gv64).max()), max(1.0, np.abs(gh64).max())
    assert np.abs(gv32 - gv64).max() <= 1e-4 * sv
    assert np.abs(gh32 - gh64).max() <= 1e-4 * sh
    _, gvf, ghf = oracle.sepconv_fwd_bwd_fast(inp, v, h, g)
    assert np.abs(gvf - gv64).max() <= 1e-4 * sv
    assert np.abs(ghf - gh64).max() <= 1e-4 * sh


def test_sepconv_unfold_torch_baseline_matches_oracle():
    p = cases.sepconv_cases()["b1c3_16x16_unit"]
    inp, v, h, g = cases.sepconv_inputs(**p)
    got = oracle.sepconv_unfold_torch(torch.from_numpy(inp), torch.from_numpy(v), torch.from_numpy(h)).numpy()
    assert np.abs(got - oracle.sepconv_forward_f64(inp, v, h)).max() <= 1e-5


def test_sepconv_oracle_matches_reference_cuda_outputs(golden_dir):
    """Pin: the reference's own .cu (compiled verbatim, run on a B200) vs the CPU restatement."""
    path = os.path.join(golden_dir, "sepconv_ref.npz")
    if not os.path.exists(path):
        pytest.skip("sepconv_ref.npz not generated yet (needs one GPU run of make_sepconv_ref_golden.py)")
    ref = np.load(path)
    for name, p in cases.sepconv_cases().items():
        if p["C"] != 3:
            continue
        inp, v, h, g = cases.sepconv_inputs(**p)
        assert np.array_equal(oracle.sepconv_forward_reforder(inp, v, h).view(np.uint32), ref[name + "_out"].view(np.uint32)), name
        assert np.array_equal(oracle.sepconv_grad_vertical_reforder(g, inp, h).view(np.uint32), ref[name + "_gv"].view(np.uint32)), name
        assert np.array_equal(oracle.sepconv_grad_horizontal_reforder(g, inp, v).view(np.uint32), ref[name + "_gh"].view(np.uint32)), name
        assert not ref[name + "_gi"].any(), "the reference leaves grad_input zero"


# ------------------------------------------------------------------ interpolation tail oracle
def test_interp_tail_restatement_matches_torch_expression():
    """oracle.interp_tail_reference against model_interp.py:90-97 evaluated with torch CPU ops
    (nn.ReplicationPad2d, torch.mean) around the reference-order sepconv."""
    r = np.random.default_rng(5)
    B, C, H, W = 2, 3, 6, 9
    i1, i2 = r.random((B, C, H, W), dtype=np.float32), r.random((B, C, H, W), dtype=np.float32)
    from sstem_restoration_b200 import synth
    k1v, k1h, k2v, k2h = (synth.unit_taps(B, 51, H, W, seed=40 + i) for i in range(4))
    pad = torch.nn.ReplicationPad2d(25)
    p1, p2 = pad(torch.from_numpy(i1)).numpy(), pad(torch.from_numpy(i2)).numpy()
    assert np.array_equal(p1, oracle._replicate_pad(i1))
    y = torch.from_numpy(oracle.sepconv_forward_reforder(p2, k2v, k2h)) + torch.from_numpy(oracle.sepconv_forward_reforder(p1, k1v, k1h))
    want = torch.mean(y, dim=1, keepdim=True).numpy()
    got = oracle.interp_tail_reference(i1, i2, k1v, k1h, k2v, k2h)
    assert got.shape == want.shape and float(np.abs(got - want).max()) <= 1.2e-7
    assert float(np.abs(got - oracle.interp_tail_f64(i1, i2, k1v, k1h, k2v, k2h)).max()) <= 1e-5


def test_interp_tail_grads_match_autograd_f64():
    r = np.random.default_rng(6)
    B, C, H, W = 1, 3, 3, 4
    i1, i2 = r.random((B, C, H, W), dtype=np.float32), r.random((B, C, H, W), dtype=np.float32)
    taps = [r.standard_normal((B, 51, H, W)).astype(np.float32) / 51 for _ in range(4)]
    g = r.standard_normal((B, 1, H, W)).astype(np.float32)
    pad = torch.nn.ReplicationPad2d(25)
    t = [torch.from_numpy(a).double().requires_grad_(True) for a in taps]
    y = _torch_sepconv_f64(pad(torch.from_numpy(i2).double()), t[2], t[3]) + _torch_sepconv_f64(pad(torch.from_numpy(i1).double()), t[0], t[1])
    torch.mean(y, dim=1, keepdim=True).backward(torch.from_numpy(g).double())
    for got, tt in zip(oracle.interp_tail_grads_f64(g, i1, i2, *taps), t):
        assert float(np.abs(got - tt.grad.numpy()).max()) <= 1e-6


# ------------------------------------------------------------------ SFF simulation (config 1) vs reference outputs
@pytest.mark.parametrize("name", list(cases.simu_sff_cases()))
def test_simu_sff_restatement_matches_reference_bitwise(golden_dir, name):
    """oracle.sff_degradation_restated / sff_noise_restated against simu_sff/simuSFF.py:96-144 run by
    tests/golden/make_golden.py with the same random.seed."""
    import hashlib
    import random
    from sstem_restoration_b200 import synth
    ref = np.load(os.path.join(golden_dir, "simu_sff_ref.npz"))
    size, index, seed = cases.simu_sff_cases()[name]
    img = synth.em_section(size, size, index)
    rng = random.Random(seed)
    deformed, flow, mask = oracle.sff_degradation_restated(img, size, rng)
    assert np.array_equal(deformed, ref[name + "_deformed"])
    assert hashlib.sha256(np.ascontiguousarray(flow).tobytes()).digest() == ref[name + "_flow_sha256"].tobytes()
    assert np.array_equal(np.packbits(mask.astype(np.uint8)), ref[name + "_mask"])
    assert np.array_equal(oracle.sff_noise_restated(deformed, size, rng), ref[name + "_noise"])


@pytest.mark.parametrize("name", list(cases.provider_degradation_cases()))
def test_provider_degradation_restatement_matches_reference_bitwise(golden_dir, name):
    """oracle.provider_degradation_restated + sff_noise_restated against the data providers' own
    `degradation` / `noise` method source executed by tests/golden/make_golden.py."""
    import hashlib
    import random
    from sstem_restoration_b200 import synth
    ref = np.load(os.path.join(golden_dir, "simu_sff_ref.npz"))
    crop, offset, index, seed, which = cases.provider_degradation_cases()[name]
    rng = random.Random(seed)
    deformed, flow2 = oracle.provider_degradation_restated(synth.em_section(crop, crop, index), crop, offset, rng,
                                                           line_width_max=50 if which == "unfolding" else 20)
    assert np.array_equal(deformed, ref[name + "_deformed"])
    assert hashlib.sha256(np.ascontiguousarray(flow2).tobytes()).digest() == ref[name + "_flow2_sha256"].tobytes()
    assert np.array_equal(oracle.sff_noise_restated(deformed, crop - 2 * offset, rng), ref[name + "_noise"])


def test_three_output_gen_flow_restatement_matches_reference_bitwise(golden_dir):
    from sstem_restoration_b200 import synth
    ref = np.load(os.path.join(golden_dir, "simu_sff_ref.npz"))
    k, b = synth.gen_line([0, 20], [64, 60])
    f1, f2, m = synth.gen_flow(64, 80, k, b, 5, 30, 0.05, two_flows=True)
    assert np.array_equal(f1.view(np.uint32), ref["gen_flow3_flow"].view(np.uint32))
    assert np.array_equal(f2.view(np.uint32), ref["gen_flow3_flow2"].view(np.uint32))
    assert np.array_equal(m.astype(np.uint8), ref["gen_flow3_mask"])


# ------------------------------------------------------------------ config 2: taps from the reference's KPN
def test_kpn_taps_fixture_and_compat_import_path(golden_dir):
    """The committed taps were predicted by the reference's own IFNet imported through compat/ (the zero-edit drop-in
    path); here: the drop-in resolves at the reference's import path, and on those raw taps the reference summation
    order is within its usual distance of fp64 (what protocol P2 measures the kernels against)."""
    import importlib
    import sys
    compat = os.path.join(os.path.dirname(golden_dir), "..", "sstem_restoration_b200", "compat")
    sys.path.insert(0, os.path.abspath(compat))
    try:
        mod = importlib.import_module("libs.sepconv.SeparableConvolution")     # model_interp.py:5
    finally:
        sys.path.pop(0)
    import sstem_restoration_b200 as pkg
    assert mod.SeparableConvolution is pkg.SeparableConvolution
    with pytest.raises(NotImplementedError):                                   # SeparableConvolution.py:47-48
        mod.SeparableConvolution.apply(torch.zeros(1, 3, 51, 51), torch.zeros(1, 51, 1, 1), torch.zeros(1, 51, 1, 1))
    p = cases.kpn_taps_case()
    taps = np.load(os.path.join(golden_dir, "kpn_taps_ref.npz"))
    assert taps["k1v"].shape == (1, 51, p["crop"], p["crop"]) and float(np.abs(taps["k2h"]).max()) > 1.0
    x = cases.kpn_frames(p)
    y0, x0, n = p["crop_y"], p["crop_x"], p["crop"]
    inp = np.ascontiguousarray(oracle._replicate_pad(x[:, 3:6])[:, :, y0:y0 + n + 50, x0:x0 + n + 50])
    ref32 = oracle.sepconv_forward_reforder(inp, taps["k2v"], taps["k2h"])
    ref64 = oracle.sepconv_forward_f64(inp, taps["k2v"], taps["k2h"])
    scale = max(1.0, float(np.abs(ref64).max()))
    assert float(np.abs(ref32 - ref64).max()) <= 1e-5 * scale


def test_warp_stitch_restatement_matches_pillow():
    """sff_scripts_fusion/inference.py:163-171 converts the warped RGB section with PIL's 'L' mode: pin the fixed-point
    luma formula of the restatement against Pillow itself (gray x3 and genuinely coloured triples)."""
    Image = pytest.importorskip("PIL.Image")
    r = np.random.default_rng(7)
    for kind in ("gray", "rgb"):
        w = r.random((3, 37, 41), dtype=np.float32)
        if kind == "gray":
            w[1:] = w[:1]
        w[:, 3:9, 5:20] = 0.003                                # warped_sff < 2: the fold line, filled from the interpolation
        interp = r.integers(0, 256, (37, 41), dtype=np.uint8)
        w8 = (w * 255).astype(np.uint8)
        pil = np.asarray(Image.fromarray(np.transpose(w8, (1, 2, 0))).convert("L"))
        mask = np.ones_like(pil, dtype=np.float32)
        mask[pil < 2] = 0
        want = (interp * (1 - mask) + pil * mask).astype(np.uint8)
        gray, stitch = oracle.warp_stitch_restated(w, interp)
        assert np.array_equal(gray, pil) and np.array_equal(stitch, want)
        if kind == "gray":
            assert np.array_equal(gray, w8[0])


def test_warp_backward_restatement_matches_reference_autograd(golden_dir):
    """The gradient formulas of oracle.warp_torch_backward_restated against autograd through the reference's own
    SpatialTransformation (tests/golden/make_warp_grad_golden.py)."""
    ref = np.load(os.path.join(golden_dir, "warp_torch_grad_ref.npz"))
    from tests.golden.make_warp_grad_golden import upstream
    for name, (mv, fl) in cases.warp_torch_cases().items():
        if name + "_gm" not in ref:
            continue
        gm, gf = oracle.warp_torch_backward_restated(mv, fl, upstream(name, mv.shape))
        assert np.abs(gm - ref[name + "_gm"]).max() <= 2e-5 * max(1.0, np.abs(ref[name + "_gm"]).max()), name
        assert np.abs(gf - ref[name + "_gf"]).max() <= 2e-5 * max(1.0, np.abs(ref[name + "_gf"]).max()), name


def test_tap_producer_restatement_matches_reference_model(golden_dir):
    """upsample(align_corners=True) -> Conv2d(51,51,3,1,1) of the reference's own IFNet tap branch, run on CPU
    (tests/golden/make_tap_producer_golden.py): indices / weights as ATen, so fp32 agrees to rounding."""
    g = np.load(os.path.join(golden_dir, "tap_producer_ref.npz"))
    up32 = oracle.upsample2x_align_corners_restated(g["x"], np.float32)
    assert np.abs(up32[:, :4] - g["up_c0_3"]).max() <= 5e-7
    y = oracle.tap_conv3x3_restated(g["x"], g["weight"], g["bias"])
    assert y.shape == g["y"].shape and np.abs(y - g["y"]).max() <= 5e-6
    # without the upsample it is a plain zero-padded cross-correlation
    r = np.random.default_rng(0)
    x, w = r.standard_normal((1, 3, 5, 6)), r.standard_normal((2, 3, 3, 3))
    want = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), padding=1).numpy()
    assert np.abs(oracle.tap_conv3x3_restated(x, w, None, upsample=False) - want).max() <= 1e-12


@pytest.mark.parametrize("h,w", [(1, 1), (1, 7), (2, 2), (3, 5), (9, 4), (16, 16), (13, 31)])
def test_upsample_restatement_matches_torch_for_odd_sizes(h, w):
    """nn.Upsample(scale_factor=2, bilinear, align_corners=True) is torch's (model_interp.py:18 only configures it): the
    restatement's float32 index / weight expressions must reproduce torch CPU for every size, degenerate ones included."""
    r = np.random.default_rng(h * 100 + w)
    x = r.standard_normal((2, 3, h, w)).astype(np.float32)
    want = torch.nn.functional.interpolate(torch.from_numpy(x), scale_factor=2, mode="bilinear", align_corners=True).numpy()
    got = oracle.upsample2x_align_corners_restated(x, np.float32)
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-6
    w9 = r.standard_normal((4, 3, 3, 3)).astype(np.float32)
    b = r.standard_normal(4).astype(np.float32)
    ref = torch.nn.functional.conv2d(torch.from_numpy(want).double(), torch.from_numpy(w9).double(), torch.from_numpy(b).double(), padding=1).numpy()
    assert np.abs(oracle.tap_conv3x3_restated(x, w9, b) - ref).max() <= 1e-5

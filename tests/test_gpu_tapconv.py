"""SURVEY 8f N2, producer side: upsample x2 (align_corners) -> Conv2d(51, 51, 3, 1, 1) as one tcgen05 TF32 kernel.

Parity bar.  The reference layer runs in TF32 (cuDNN, torch's default allow_tf32): both operands carry 10 explicit
mantissa bits, products are accumulated in fp32.  This kernel rounds operands to nearest, so per product the relative
error is at most 2^-10 (+ second order); the bound checked is

    |got - fp64 oracle| <= 1.05 * 2^-10 * (conv(|up(x)|, |w|) + |bias|) + 1e-6

and in practice the error is several times below it (random signs).  Between the two output layouts and from run to run
the results must be bit-identical."""
import os

import numpy as np
import pytest
import torch

import oracle
import sstem_restoration_b200 as pkg

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _run(x, w, b, upsample, tiled=False):
    tx, tw = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
    tb = torch.from_numpy(b).cuda() if b is not None else None
    return pkg.tap_conv3x3(tx, pkg.pack_tap_conv_weight(tw), tb, upsample=upsample, tiled=tiled)


def _check(got, x, w, b, upsample):
    want = oracle.tap_conv3x3_restated(x, w, b, upsample)
    bound = oracle.tap_conv3x3_restated(np.abs(x), np.abs(w), None if b is None else np.abs(b), upsample)
    err = np.abs(got.astype(np.float64) - want)
    assert got.shape == want.shape
    assert np.isfinite(got).all()
    assert (err <= 1.05 * 2.0 ** -10 * bound + 1e-6).all(), float((err / (2.0 ** -10 * bound + 1e-9)).max())
    return float(err.max()), float((err / (2.0 ** -10 * bound + 1e-12)).max())


def test_reference_model_layer(golden_dir=GOLDEN):
    g = np.load(os.path.join(golden_dir, "tap_producer_ref.npz"))
    got = _run(g["x"], g["weight"], g["bias"], True).cpu().numpy()
    _check(got, g["x"], g["weight"], g["bias"], True)
    # and against the reference's own fp32 output: TF32-sized differences only
    assert np.abs(got - g["y"]).max() <= 2.0 ** -9 * np.abs(g["y"]).max()


@pytest.mark.parametrize("B,cin,cout,h,w,ups", [
    (1, 51, 51, 8, 4, True), (2, 51, 51, 13, 9, True), (1, 51, 51, 16, 8, False), (1, 51, 51, 37, 29, False),
    (1, 8, 16, 5, 7, True), (1, 3, 5, 9, 6, False), (1, 52, 64, 6, 6, True), (1, 52, 64, 6, 5, True), (2, 17, 33, 20, 11, False),
    (1, 51, 51, 1, 1, True), (1, 51, 51, 1, 1, False), (1, 4, 8, 2, 33, True),
])
def test_small_and_ragged_shapes(B, cin, cout, h, w, ups):
    r = np.random.default_rng(B * 1000 + cin * 7 + cout + h * 13 + w)
    x = r.standard_normal((B, cin, h, w)).astype(np.float32)
    wt = (r.standard_normal((cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
    b = r.standard_normal(cout).astype(np.float32)
    _check(_run(x, wt, b, ups).cpu().numpy(), x, wt, b, ups)
    _check(_run(x, wt, None, ups).cpu().numpy(), x, wt, None, ups)


def test_many_tiles_per_cta_both_layouts_and_determinism():
    """256 x 256 output = 512 tiles over 148 CTAs: every barrier runs through both phases several times."""
    r = np.random.default_rng(5)
    x = np.maximum(r.standard_normal((1, 51, 128, 128)), 0).astype(np.float32)      # post-ReLU activations
    wt = (r.standard_normal((51, 51, 3, 3)) / np.sqrt(459)).astype(np.float32)
    b = (0.1 * r.standard_normal(51)).astype(np.float32)
    a = _run(x, wt, b, True)
    _check(a.cpu().numpy(), x, wt, b, True)
    assert torch.equal(a, _run(x, wt, b, True))
    t = _run(x, wt, b, True, tiled=True)
    assert torch.equal(t, pkg.taps_to_tiled(a))


@pytest.mark.parametrize("H,W", [(40, 24), (37, 29)])
def test_tiled_output_ragged_matches_layout_conversion(H, W):
    r = np.random.default_rng(H + W)
    x = r.standard_normal((2, 51, H, W)).astype(np.float32)
    wt = (r.standard_normal((51, 51, 3, 3)) / 20).astype(np.float32)
    a = _run(x, wt, None, False)
    t = _run(x, wt, None, False, tiled=True)
    assert torch.equal(t, pkg.taps_to_tiled(a))


def test_producer_feeds_consumer_without_nchw_taps():
    """The N2 chain: half-resolution activations -> tile-major taps -> sepconv, no [B,51,H,W] tensor in between;
    identical to routing the same taps through the NCHW operator."""
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(12)
    B, H, W = 1, 256, 256
    act = [torch.relu(torch.randn((B, 51, H // 2, W // 2), device=dev, generator=gen)) for _ in range(2)]
    mods = [pkg.ModuleTapProducer(tiled=True).to(dev) for _ in range(2)]
    frame = torch.rand((B, 3, H + 50, W + 50), device=dev, generator=gen)
    vt, ht = mods[0](act[0]), mods[1](act[1])
    got = pkg.sepconv_forward_tiled(frame, vt, ht)
    v = pkg.tap_conv3x3(act[0], mods[0]._packed, mods[0].bias.detach())
    h = pkg.tap_conv3x3(act[1], mods[1]._packed, mods[1].bias.detach())
    want = pkg.SeparableConvolution.apply(frame, v, h)
    assert torch.equal(got, want)
    # the module tracks weight updates
    with torch.no_grad():
        mods[0].weight.mul_(2.0)
    assert not torch.equal(mods[0](act[0]), vt)


def test_module_loads_reference_conv_state_dict():
    conv = torch.nn.Conv2d(51, 51, 3, 1, 1)
    m = pkg.ModuleTapProducer()
    m.load_state_dict(conv.state_dict())
    m = m.cuda()
    x = torch.rand((1, 51, 8, 8), device="cuda")
    ref = conv.cuda()(torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True))
    assert (m(x) - ref).abs().max() <= 2.0 ** -9 * ref.abs().max()


def test_large_image_against_fp32_library_convolution():
    """1024^2 output (8192 tiles, 55 per CTA, image borders on all sides): the fp64 oracle is too slow here, so the comparison
    is against torch's own fp32 (TF32 disabled) interpolate + conv2d on the same GPU, with the same TF32 bound."""
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(77)
    x = torch.relu(torch.randn((1, 51, 512, 512), device=dev, generator=gen))
    w = torch.randn((51, 51, 3, 3), device=dev, generator=gen) / 21.4
    b = 0.1 * torch.randn(51, device=dev, generator=gen)
    got = pkg.tap_conv3x3(x, pkg.pack_tap_conv_weight(w), b)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        up = torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        want = torch.nn.functional.conv2d(up, w, b, padding=1)
        bound = torch.nn.functional.conv2d(up.abs(), w.abs(), b.abs(), padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    err = (got - want).abs()
    assert bool((err <= 1.05 * 2.0 ** -10 * bound + 1e-5).all()), float((err / (2.0 ** -10 * bound + 1e-9)).max())
    # typical error: well inside the bound (random signs), and no systematic offset
    assert float(err.mean()) <= 0.1 * float((2.0 ** -10 * bound).mean())
    assert abs(float((got - want).mean())) <= 1e-5
    tiled = pkg.tap_conv3x3(x, pkg.pack_tap_conv_weight(w), b, tiled=True)
    assert torch.equal(tiled, pkg.taps_to_tiled(got))


def test_forward_only_is_loud_under_autograd():
    x = torch.rand((1, 51, 8, 8), device="cuda", requires_grad=True)
    m = pkg.ModuleTapProducer().cuda()
    with pytest.raises(NotImplementedError):
        m(x)
    with torch.no_grad():
        assert m(x).shape == (1, 51, 16, 16)
    assert m(x.detach()).requires_grad is False

"""GPU parity tests for the fused interpolation tail: one launch for
model_interp.py:90-97 = 2x ReplicationPad2d(25) + 2x SeparableConvolution + add + channel mean.

Tolerance: 1e-5 max-abs on unit-scale inputs against the reference-order oracle (the fused
kernel takes the channel mean before the convolution; the convolution is linear in the image,
so the two agree to fp32 rounding), plus bit-exact size-independent properties at full size.
"""
import numpy as np
import pytest
import torch

import oracle
from sstem_restoration_b200 import synth

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _pkg():
    import sstem_restoration_b200 as pkg
    return pkg


def _inputs(B, C, H, W, seed, gray=False):
    r = np.random.default_rng(seed)
    if gray:
        i1 = np.repeat(r.random((B, 1, H, W), dtype=np.float32), C, axis=1)
        i2 = np.repeat(r.random((B, 1, H, W), dtype=np.float32), C, axis=1)
    else:
        i1 = r.random((B, C, H, W), dtype=np.float32)
        i2 = r.random((B, C, H, W), dtype=np.float32)
    taps = [synth.unit_taps(B, 51, H, W, seed=seed + 10 + i) for i in range(4)]
    return (i1, i2, *taps)


def _cuda(*arrs):
    return tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs)


def _unfused(pkg, i1, i2, k1v, k1h, k2v, k2h):
    pad = torch.nn.ReplicationPad2d(25)
    y = pkg.SeparableConvolution.apply(pad(i2).contiguous(), k2v, k2h) + pkg.SeparableConvolution.apply(pad(i1).contiguous(), k1v, k1h)
    return torch.mean(y, dim=1, keepdim=True)


@pytest.mark.parametrize("B,C,H,W", [(1, 3, 64, 64), (2, 3, 33, 65), (1, 1, 40, 72), (1, 3, 8, 129), (1, 2, 5, 7),
                                      (3, 3, 1, 1), (1, 4, 17, 100)])
def test_forward_parity_general_channels(B, C, H, W):
    arrs = _inputs(B, C, H, W, seed=500 + H + W)
    got = _pkg().interpolation_tail(*_cuda(*arrs)).cpu().numpy()
    ref = oracle.interp_tail_reference(*arrs)
    assert got.shape == ref.shape == (B, 1, H, W)
    assert float(np.abs(got.astype(np.float64) - ref).max()) <= TOL
    ref64 = oracle.interp_tail_f64(*arrs)
    assert float(np.abs(got - ref64).max()) <= max(1.25 * float(np.abs(ref - ref64).max()), 2e-6)


@pytest.mark.parametrize("mode", ["assert", "detect"])
def test_forward_gray_replicated(mode):
    pkg = _pkg()
    arrs = _inputs(2, 3, 40, 48, seed=77, gray=True)
    t = _cuda(*arrs)
    pkg.set_gray_replicated(mode)
    try:
        got = pkg.interpolation_tail(*t).cpu().numpy()
    finally:
        pkg.set_gray_replicated("off")
    ref = oracle.interp_tail_reference(*arrs)
    assert float(np.abs(got.astype(np.float64) - ref).max()) <= TOL
    unf = _unfused(pkg, *t).cpu().numpy()
    assert float(np.abs(got - unf).max()) <= TOL


def test_detect_mode_falls_back_when_planes_differ():
    pkg = _pkg()
    arrs = _inputs(1, 3, 24, 40, seed=78)
    pkg.set_gray_replicated("detect")
    try:
        got = pkg.interpolation_tail(*_cuda(*arrs)).cpu().numpy()
    finally:
        pkg.set_gray_replicated("off")
    assert float(np.abs(got - oracle.interp_tail_reference(*arrs)).max()) <= TOL


def test_frames_as_views_of_the_network_input():
    """model_interp.py:56-57: i1 = x[:, :3], i2 = x[:, 3:6] -- taken without a copy."""
    pkg = _pkg()
    i1, i2, k1v, k1h, k2v, k2h = _inputs(2, 3, 32, 36, seed=79)
    x = torch.from_numpy(np.concatenate([i1, i2], axis=1)).cuda()
    taps = _cuda(k1v, k1h, k2v, k2h)
    got = pkg.interpolation_tail(x[:, :3], x[:, 3:6], *taps).cpu().numpy()
    assert float(np.abs(got - oracle.interp_tail_reference(i1, i2, k1v, k1h, k2v, k2h)).max()) <= TOL


@pytest.mark.parametrize("B,C,H,W,gray", [(1, 3, 16, 32, False), (2, 3, 9, 13, False), (1, 3, 20, 24, True), (1, 1, 12, 40, False)])
def test_backward_parity(B, C, H, W, gray):
    pkg = _pkg()
    arrs = _inputs(B, C, H, W, seed=900 + H, gray=gray)
    g = np.random.default_rng(99).standard_normal((B, 1, H, W)).astype(np.float32)
    i1, i2, *taps = _cuda(*arrs)
    taps = [t.requires_grad_(True) for t in taps]
    pkg.set_gray_replicated("assert" if gray else "off")
    try:
        out = pkg.interpolation_tail(i1, i2, *taps)
        out.backward(torch.from_numpy(g).cuda())
    finally:
        pkg.set_gray_replicated("off")
    ref = oracle.interp_tail_grads_f64(g, *arrs)
    for name, t, r in zip(("k1v", "k1h", "k2v", "k2h"), taps, ref):
        err = float(np.abs(t.grad.cpu().numpy() - r).max())
        assert err <= TOL * max(1.0, float(np.abs(r).max())), (name, err)
    # and against autograd through the unfused expression on the same device
    taps2 = [t.detach().clone().requires_grad_(True) for t in taps]
    _unfused(pkg, i1, i2, *taps2).backward(torch.from_numpy(g).cuda())
    for t, t2 in zip(taps, taps2):
        assert float((t.grad - t2.grad).abs().max()) <= TOL * max(1.0, float(t2.grad.abs().max()))


def test_backward_honours_needs_input_grad():
    pkg = _pkg()
    i1, i2, k1v, k1h, k2v, k2h = _cuda(*_inputs(1, 3, 8, 32, seed=5))
    k2h.requires_grad_(True)
    pkg.interpolation_tail(i1, i2, k1v, k1h, k2v, k2h).sum().backward()
    assert k2h.grad is not None and k1v.grad is None and k1h.grad is None and k2v.grad is None
    i1.requires_grad_(True)
    with pytest.raises(NotImplementedError):
        pkg.interpolation_tail(i1, i2, k1v, k1h, k2v, k2h).sum().backward()


def test_reference_error_behaviour():
    pkg = _pkg()
    arrs = _inputs(1, 3, 8, 8, seed=6)
    with pytest.raises(NotImplementedError):          # SeparableConvolution.py:47-48
        pkg.interpolation_tail(*(torch.from_numpy(a) for a in arrs))
    i1, i2, k1v, k1h, k2v, k2h = _cuda(*arrs)
    with pytest.raises(AssertionError):               # SeparableConvolution.py:31: 51 taps
        pkg.interpolation_tail(i1, i2, k1v[:, :49].contiguous(), k1h, k2v, k2h)


@pytest.mark.parametrize("H,W", [(2048, 2048), (512, 516)])
def test_full_size_one_hot_taps_gather_exactly(H, W):
    """One-hot taps turn each sepconv into a gather from the replicate-padded frame, so the fused
    result must equal frame2[clamp] + frame1[clamp] bit for bit (gray sections, as the callers feed)."""
    pkg = _pkg()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(3)
    i1 = torch.rand((1, 1, H, W), device=dev, generator=gen).expand(1, 3, H, W).contiguous()
    i2 = torch.rand((1, 1, H, W), device=dev, generator=gen).expand(1, 3, H, W).contiguous()
    fy = torch.randint(0, 51, (4, 1, 1, H, W), device=dev, generator=gen)
    taps = [torch.zeros((1, 51, H, W), device=dev).scatter_(1, fy[i], 1.0) for i in range(4)]
    k1v, k1h, k2v, k2h = taps
    yy = torch.arange(H, device=dev).view(H, 1)
    xx = torch.arange(W, device=dev).view(1, W)

    def gather(img, fv, fh):
        sy = (yy + fv[0, 0] - 25).clamp(0, H - 1)
        sx = (xx + fh[0, 0] - 25).clamp(0, W - 1)
        return img[0, 0][sy, sx]

    want = gather(i2, fy[2], fy[3]) + gather(i1, fy[0], fy[1])
    pkg.set_gray_replicated("assert")
    try:
        got = pkg.interpolation_tail(i1, i2, k1v, k1h, k2v, k2h)
    finally:
        pkg.set_gray_replicated("off")
    assert torch.equal(got[0, 0], want)
    # general path: the channel sum of three equal planes is 3x (exact up to one rounding), then / 3
    got3 = pkg.interpolation_tail(i1, i2, k1v, k1h, k2v, k2h)
    assert float((got3[0, 0] - want).abs().max()) <= 2e-7 * 2

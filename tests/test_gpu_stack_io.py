"""GPU parity tests for the stack pre/post-processing kernels: bit-exact
against numpy restatements of sff_scripts_interp/inference.py:69-88."""
import numpy as np
import pytest
import torch

import oracle
import sstem_restoration_b200 as pkg
from sstem_restoration_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,pad", [(64, 64, 0), (37, 53, 5), (1, 1, 3), (130, 258, 16), (256, 256, 25)])
def test_sections_to_input_bit_exact(H, W, pad):
    r = np.random.default_rng(H * W + pad)
    a, b = r.integers(0, 256, (H, W), dtype=np.uint8), r.integers(0, 256, (H, W), dtype=np.uint8)
    got = pkg.sections_to_input(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), pad)
    want = oracle.sections_to_input_restated(a, b, pad)
    assert got.shape == want.shape and got.dtype == torch.float32
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))
    got_host = pkg.sections_to_input(a, b, pad)          # numpy in: uploaded, result stays on the device
    assert got_host.is_cuda and torch.equal(got_host, got)


def test_sections_to_input_batched_matches_per_section():
    r = np.random.default_rng(3)
    a, b = r.integers(0, 256, (3, 40, 44), dtype=np.uint8), r.integers(0, 256, (3, 40, 44), dtype=np.uint8)
    got = pkg.sections_to_input(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 7).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i:i + 1], oracle.sections_to_input_restated(a[i], b[i], 7))


@pytest.mark.parametrize("H,W,pad", [(64, 64, 0), (37, 53, 5), (1, 1, 3), (130, 259, 16)])
def test_prediction_to_uint8_bit_exact(H, W, pad):
    r = np.random.default_rng(H + W)
    pred = r.random((1, 1, H + 2 * pad, W + 2 * pad), dtype=np.float32)
    pred.flat[:: 7] = np.float32(1.0)                     # 255 exactly
    pred.flat[:: 11] = np.float32(0.0)
    got = pkg.prediction_to_uint8(torch.from_numpy(pred).cuda(), pad).cpu().numpy()
    want = oracle.prediction_to_uint8_restated(pred, pad).reshape(1, H, W)
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    assert np.array_equal(pkg.prediction_to_uint8(pred[0, 0], pad), want[0])     # numpy [H,W] in -> numpy out


def test_round_trip_full_size_4096():
    """uint8 -> /255 -> *255 -> uint8 is the identity for every byte value (size-independent property)."""
    sec = torch.from_numpy(synth.em_section(512, 512, 1)).cuda().repeat(8, 8).contiguous()
    x = pkg.sections_to_input(sec, sec.flip(0), 16)
    assert x.shape == (1, 6, 4096 + 32, 4096 + 32)
    assert float(x[:, :, :16].abs().max()) == 0.0 and float(x[:, :, :, -16:].abs().max()) == 0.0
    assert torch.equal(x[0, 0], x[0, 2]) and torch.equal(x[0, 3], x[0, 5])
    back = pkg.prediction_to_uint8(x[:, :1].contiguous(), 16)
    want = (sec.float() / 255.0 * 255).to(torch.uint8)
    assert torch.equal(back[0], want)


def test_type_errors():
    with pytest.raises(TypeError):
        pkg.sections_to_input(torch.zeros((4, 4), device="cuda"), torch.zeros((4, 4), device="cuda"))
    with pytest.raises(TypeError):
        pkg.prediction_to_uint8(torch.zeros((4, 4), dtype=torch.uint8, device="cuda"))

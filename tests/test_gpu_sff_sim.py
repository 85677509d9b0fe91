"""GPU parity tests for the GPU-resident SFF simulation (BASELINE config 1):
sstem_sff_degrade / sstem_sff_contrast through the host mirror of simu_sff/simuSFF.py.

Bar: bit-exact (uint8 images, float32 flow bits, mask, zero count) against the reference's own
outputs (tests/golden/simu_sff_ref.npz) and against the numpy oracle."""
import hashlib
import os
import random

import numpy as np
import pytest
import torch

import oracle
from sstem_restoration_b200 import sff_sim, synth
from tests.golden import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(cases.simu_sff_cases()))
@pytest.mark.parametrize("host", [False, True])
def test_degradation_and_noise_match_reference_outputs(golden_dir, name, host):
    ref = np.load(os.path.join(golden_dir, "simu_sff_ref.npz"))
    size, index, seed = cases.simu_sff_cases()[name]
    img = synth.em_section(size, size, index)
    rng = random.Random(seed)
    src = img if host else torch.from_numpy(img).cuda()
    deformed, flow, mask = sff_sim.degradation(src, size, rng=rng)
    out = sff_sim.noise(deformed, size, rng=rng)
    if not host:
        deformed, flow, mask, out = (t.cpu().numpy() for t in (deformed, flow, mask, out))
    assert deformed.dtype == np.uint8 and np.array_equal(deformed, ref[name + "_deformed"])
    assert hashlib.sha256(np.ascontiguousarray(flow).tobytes()).digest() == ref[name + "_flow_sha256"].tobytes()
    assert np.array_equal(np.packbits(mask.astype(np.uint8)), ref[name + "_mask"])
    assert np.array_equal(out, ref[name + "_noise"])


def test_simu_sff_pipeline_matches_oracle_with_crop():
    """SimuSFF's crop + degradation + noise (simuSFF.py:14-30) on a section larger than the patch."""
    img = synth.em_section(384, 400, 9)
    got, flow, mask = sff_sim.simu_sff(torch.from_numpy(img).cuda(), 256, rng=random.Random(31))
    rng = random.Random(31)
    i, j = rng.randint(0, 384 - 256), rng.randint(0, 400 - 256)
    d, f, m = oracle.sff_degradation_restated(img[i:i + 256, j:j + 256], 256, rng)
    want = oracle.sff_noise_restated(d, 256, rng)
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(flow.cpu().numpy().view(np.uint32), f.view(np.uint32))
    assert np.array_equal(mask.cpu().numpy(), m.astype(np.uint8))


@pytest.mark.parametrize("name", list(cases.gen_flow_cases()))
def test_kernel_matches_gen_flow_and_image_warp(name):
    h, w, p1, p2, lw, fw, dk = cases.gen_flow_cases()[name]
    k, b = synth.gen_line(p1, p2)
    img = np.random.default_rng(h * w).integers(0, 256, (h, w), dtype=np.uint8)
    out, flow, mask, stats = sff_sim.gen_flow_warp(torch.from_numpy(img).cuda()[None], [sff_sim.fold_line_params(k, b, lw, fw, dk)])
    rflow, rmask = synth.gen_flow(h, w, k, b, lw, fw, dk)
    want = (oracle.image_warp_restated(img, rflow) * rmask).astype(np.uint8)
    assert np.array_equal(flow[0].cpu().numpy().view(np.uint32), rflow.view(np.uint32))
    assert np.array_equal(mask[0].cpu().numpy(), rmask.astype(np.uint8))
    assert np.array_equal(out[0].cpu().numpy(), want)
    assert stats.cpu().tolist() == [[int((want == 0).sum()), int(want.sum(dtype=np.int64))]]


def test_ragged_sizes_batches_and_optional_outputs():
    rng = np.random.default_rng(8)
    for (B, H, W) in [(3, 37, 53), (2, 1, 7), (1, 130, 258), (2, 64, 66)]:
        imgs = rng.integers(0, 256, (B, H, W), dtype=np.uint8)
        lines = [synth.gen_line([0, 3 + 5 * i], [H, W - 2 - 3 * i]) for i in range(B)]
        prm = [(k, b, 3 + i, 9 + 4 * i, 0.02 * (i + 1)) for i, (k, b) in enumerate(lines)]
        out, flow, mask, stats = sff_sim.gen_flow_warp(torch.from_numpy(imgs).cuda(), [sff_sim.fold_line_params(*p) for p in prm],
                                                       want_flow=False, want_mask=False)
        assert flow is None and mask is None
        for i, (k, b, lw, fw, dk) in enumerate(prm):
            rflow, rmask = synth.gen_flow(H, W, k, b, lw, fw, dk)
            want = (oracle.image_warp_restated(imgs[i], rflow) * rmask).astype(np.uint8)
            assert np.array_equal(out[i].cpu().numpy(), want), (B, H, W, i)
            assert stats[i].cpu().tolist() == [int((want == 0).sum()), int(want.sum(dtype=np.int64))]


def test_full_size_properties_4096():
    """4096^2 (config-5 section size): where the synthesised flow is exactly zero the section is
    copied unchanged, masked pixels are zero, and the statistics equal recounts of the output."""
    H = W = 4096
    img = torch.from_numpy(synth.em_section(512, 512, 2)).cuda().repeat(8, 8).contiguous()
    k, b = synth.gen_line([0, 1500], [H, 2600])
    out, flow, mask, stats = sff_sim.gen_flow_warp(img[None], [sff_sim.fold_line_params(k, b, 12, 60, 0.05)])
    out, flow, mask = out[0], flow[0], mask[0]
    still = (flow[..., 0] == 0) & (flow[..., 1] == 0) & (mask == 1)   # (the line itself has zero flow but is masked)
    assert int(still.sum()) > H * W // 4
    assert torch.equal(out[still], img[still])
    assert int(out[mask == 0].max()) == 0 and int((mask == 0).sum()) > 0
    assert stats.cpu().tolist() == [[int((out == 0).sum()), int(out.sum(dtype=torch.int64))]]


def test_rejects_non_uint8_and_bad_rank():
    with pytest.raises(TypeError):
        sff_sim.degradation(torch.zeros((256, 256), device="cuda"), 256)
    with pytest.raises(ValueError):
        sff_sim.degradation(torch.zeros((2, 2, 2, 2), dtype=torch.uint8, device="cuda"), 256)


@pytest.mark.parametrize("name", list(cases.provider_degradation_cases()))
def test_provider_degradation_matches_reference_outputs(golden_dir, name):
    """The data providers' degradation + noise (data_provider.py:180-259) on the GPU vs the reference's
    own method source run with the same random.seed."""
    ref = np.load(os.path.join(golden_dir, "simu_sff_ref.npz"))
    crop, offset, index, seed, which = cases.provider_degradation_cases()[name]
    rng = random.Random(seed)
    img = torch.from_numpy(synth.em_section(crop, crop, index)).cuda()
    deformed, flow2 = sff_sim.provider_degradation(img, crop, offset, rng=rng, line_width_max=50 if which == "unfolding" else 20)
    out = sff_sim.noise(deformed, crop - 2 * offset, rng=rng)
    assert np.array_equal(deformed.cpu().numpy(), ref[name + "_deformed"])
    assert hashlib.sha256(np.ascontiguousarray(flow2.cpu().numpy()).tobytes()).digest() == ref[name + "_flow2_sha256"].tobytes()
    assert np.array_equal(out.cpu().numpy(), ref[name + "_noise"])


def test_gen_flow_drop_in_bit_exact(golden_dir):
    ref = np.load(os.path.join(golden_dir, "gen_flow_ref.npz"))
    for name, (h, w, p1, p2, lw, fw, dk) in cases.gen_flow_cases().items():
        k, b = sff_sim.gen_line(p1, p2)
        flow, mask = sff_sim.gen_flow(h, w, k, b, lw, fw, dk)
        assert flow.dtype == np.float32 and mask.dtype == np.float64
        assert np.array_equal(flow.view(np.uint32), ref[name + "_flow"].view(np.uint32)), name
        assert np.array_equal(mask.astype(np.uint8), ref[name + "_mask"]), name
    ref3 = np.load(os.path.join(golden_dir, "simu_sff_ref.npz"))
    k, b = sff_sim.gen_line([0, 20], [64, 60])
    f1, f2, m = sff_sim.gen_flow(64, 80, k, b, 5, 30, 0.05, two_flows=True)
    assert np.array_equal(f1.view(np.uint32), ref3["gen_flow3_flow"].view(np.uint32))
    assert np.array_equal(f2.view(np.uint32), ref3["gen_flow3_flow2"].view(np.uint32))
    assert np.array_equal(m.astype(np.uint8), ref3["gen_flow3_mask"])


@pytest.mark.parametrize("seed,with_noise", [(0, True), (7, True), (2, False)])
def test_provider_batch_equals_sequential_reference(seed, with_noise):
    """One degrade launch + one contrast launch per pass for the whole batch, same draws and outputs as B sequential
    Provider.degradation / Provider.noise calls (several samples of these seeds need 2-5 attempts)."""
    crop, offset, det, B = 448, 96, 256, 5
    imgs = np.stack([synth.em_section(crop, crop, 20 + i) for i in range(B)])
    rng = random.Random(seed)
    sff, flow2 = sff_sim.provider_batch(torch.from_numpy(imgs).cuda(), crop, offset, rng=rng, with_noise=with_noise)
    assert sff.shape == (B, det, det) and flow2.shape == (B, det, det, 2)
    ref = random.Random(seed)
    for i in range(B):
        d, f2 = oracle.provider_degradation_restated(imgs[i], crop, offset, ref, 50)
        want = oracle.sff_noise_restated(d, det, ref) if with_noise else d
        assert np.array_equal(sff[i].cpu().numpy(), want), (seed, i)
        assert np.array_equal(flow2[i].cpu().numpy().view(np.uint32), f2.view(np.uint32)), (seed, i)
    assert rng.random() == ref.random()

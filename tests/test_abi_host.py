"""CPU: the C-ABI library loads and exports what include/*.h declares; host-side logic."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sstem_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sstem_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) >= 8
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/sstem_b200.h but not exported"
    from sstem_restoration_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_abi_version_and_error_strings(built_lib):
    from sstem_restoration_b200 import _lib
    lib = _lib.load()
    assert lib.sstem_abi_version() == _lib.ABI_VERSION
    assert lib.sstem_error_string(0) == b"success"
    for code in (-1, -2, -3, -4, -5):
        assert lib.sstem_error_string(code).startswith(b"sstem:")
    assert lib.sstem_launch_count() >= 0


def test_argument_errors_do_not_need_a_gpu(built_lib):
    from sstem_restoration_b200 import _lib
    lib = _lib.load()
    assert lib.sstem_sepconv_forward(None, None, None, None, 1, 1, 1, 1, 51, 0, None) == -1
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    assert lib.sstem_sepconv_forward(p, p, p, p, 0, 3, 4, 4, 51, 0, None) == -2
    assert lib.sstem_sepconv_forward(p, p, p, p, 1, 3, 4, 4, 65, 0, None) == -2
    assert lib.sstem_sepconv_forward(p, p, p, p, 1, 3, 4, 4, 51, 8, None) == -5
    assert lib.sstem_sepconv_forward(p + 1, p, p, p, 1, 3, 4, 4, 51, 0, None) == -3
    assert lib.sstem_sepconv_backward(p, p, p, p, None, None, None, 1, 3, 4, 4, 51, 0, None) == -1
    strides = (ctypes.c_int64 * 4)(32, 8, 2, 1)
    assert lib.sstem_warp_forward(p, p, strides, p, 1, 1, 2, 2, 7, None) == -5
    assert lib.sstem_image_warp(p, 9, p, p, None, 1, 2, 2, 1, 0, None) == -5
    with pytest.raises(_lib.SstemError):
        _lib.check(-2, "x")


def test_sepconv_asserts_and_cpu_error_mirror_reference(built_lib):
    """libs/sepconv/SeparableConvolution.py:29-35 asserts, :47-48 NotImplementedError on CPU."""
    from sstem_restoration_b200 import SeparableConvolution, FunctionSepconv, ModuleSepconv
    inp = torch.zeros(1, 3, 60, 60)
    v = torch.zeros(1, 51, 10, 10)
    with pytest.raises(NotImplementedError):
        SeparableConvolution.apply(inp, v, v)
    with pytest.raises(AssertionError):
        SeparableConvolution.apply(torch.zeros(1, 3, 61, 60), v, v)        # height mismatch
    with pytest.raises(AssertionError):
        SeparableConvolution.apply(torch.zeros(1, 3, 58, 58), torch.zeros(1, 49, 10, 10), torch.zeros(1, 49, 10, 10))  # K != 51
    with pytest.raises(AssertionError):
        SeparableConvolution.apply(inp.transpose(2, 3), v, v) if False else SeparableConvolution.apply(inp[:, :, :, ::1].permute(0, 1, 3, 2), v, v)
    # generic-K variant accepts 49 taps but still refuses CPU tensors
    with pytest.raises(NotImplementedError):
        FunctionSepconv(torch.zeros(1, 3, 58, 58), torch.zeros(1, 49, 10, 10), torch.zeros(1, 49, 10, 10))
    with pytest.raises(NotImplementedError):
        ModuleSepconv()(torch.zeros(1, 3, 58, 58), torch.zeros(1, 49, 10, 10), torch.zeros(1, 49, 10, 10))


def test_compat_import_paths(built_lib):
    """The reference's dotted import paths resolve to the sm_100a implementation."""
    compat = os.path.join(ROOT, "sstem_restoration_b200", "compat")
    sys.path.insert(0, compat)
    try:
        for m in [k for k in sys.modules if k == "libs" or k.startswith("libs.") or k == "utils" or k.startswith("utils.") or k == "model" or k.startswith("model.")]:
            del sys.modules[m]
        from libs.sepconv.SeparableConvolution import SeparableConvolution as S1
        import sstem_restoration_b200 as pkg
        assert S1 is pkg.SeparableConvolution
        import importlib.util
        for rel, attr in (("utils/image_warp_torch.py", "SpatialTransformation"), ("utils/image_warp.py", "image_warp"),
                          ("model/sepconv.py", "FunctionSepconv")):
            spec = importlib.util.spec_from_file_location("compat_" + attr, os.path.join(compat, rel))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            assert getattr(mod, attr) is getattr(pkg, attr) or getattr(mod, attr) is getattr(pkg.sepconv, attr, None)
    finally:
        sys.path.remove(compat)


def test_warp_refuses_to_run_without_cuda(built_lib):
    from sstem_restoration_b200 import SpatialTransformation, image_warp, SstemError
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(SstemError):
        SpatialTransformation()(torch.zeros(1, 1, 4, 4), torch.zeros(1, 4, 4, 2))
    with pytest.raises(SstemError):
        image_warp(np.zeros((4, 4), np.uint8), np.zeros((4, 4, 2), np.float32))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from sstem_restoration_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.SstemError):
        _lib.load()


def test_shard_ranges_partition_units():
    from sstem_restoration_b200 import shard
    for n in (0, 1, 7, 98, 100):
        for ws in (1, 2, 4, 8):
            spans = [shard.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == shard.max_units_per_rank(n, ws)
    assert shard.stack_targets(100)[0] == (0, 1, 2) and len(shard.stack_targets(100)) == 98
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_synth_inputs_are_deterministic_and_shaped():
    from sstem_restoration_b200 import synth
    a, b = synth.em_section(96, 80, 3), synth.em_section(96, 80, 3)
    assert a.dtype == np.uint8 and a.shape == (96, 80) and np.array_equal(a, b)
    assert not np.array_equal(a, synth.em_section(96, 80, 4))
    x = synth.section_to_input(a)
    assert x.shape == (3, 146, 130) and x.dtype == np.float32 and np.array_equal(x[0], x[2])
    assert np.array_equal(x[0, 25:-25, 25:-25], a.astype(np.float32) / np.float32(255))
    t = synth.unit_taps(2, 51, 8, 8)
    assert t.shape == (2, 51, 8, 8) and np.allclose(t.sum(1), 1, atol=1e-5)
    f, m = synth.random_fold_flow(128, 128)
    assert f.shape == (128, 128, 2) and f.dtype == np.float32 and set(np.unique(m)) <= {0.0, 1.0}


def test_argument_errors_of_the_section_8f_entry_points(built_lib):
    """Fused tail, SFF simulation and stack I/O: argument checks run before any CUDA call."""
    from sstem_restoration_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(256)
    p = ctypes.addressof(buf)
    taps = [p] * 4
    assert lib.sstem_interp_tail_forward(None, p, 48, *taps, p, 1, 3, 4, 4, 51, 0, None) == -1
    assert lib.sstem_interp_tail_forward(p, p, 48, *taps, p, 1, 3, 4, 4, 49, 0, None) == -2       # 51 taps only
    assert lib.sstem_interp_tail_forward(p, p, 47, *taps, p, 1, 3, 4, 4, 51, 0, None) == -2       # batch stride < C*H*W
    assert lib.sstem_interp_tail_forward(p, p, 48, *taps, p, 1, 3, 4, 4, 51, 1, None) == -5       # STRICT_ORDER not offered
    assert lib.sstem_interp_tail_forward(p, p + 2, 48, *taps, p, 1, 3, 4, 4, 51, 0, None) == -3
    assert lib.sstem_interp_tail_backward(p, p, p, 48, *taps, None, None, None, None, 1, 3, 4, 4, 51, 0, None) == -1
    assert lib.sstem_sff_degrade(None, p, p, None, None, None, p, 1, 4, 4, 0, None) == -1
    assert lib.sstem_sff_degrade(p, p, p, None, None, None, p, 1, 4, 4, 2, None) == -2           # border swallows the image
    assert lib.sstem_sff_degrade(p, p + 4, p, None, None, None, p, 1, 4, 4, 0, None) == -3       # params not 8-byte aligned
    assert lib.sstem_sff_contrast(p, p, p, 1, 4, 4, 0, 4, None) == -2
    assert lib.sstem_sections_to_input(None, p, p, 1, 4, 4, 0, None) == -1                        # (section_next may be NULL)
    assert lib.sstem_warp_stitch_u8(p, None, p, p, 1, 3, 4, 4, None) == -1
    assert lib.sstem_warp_stitch_u8(p, p, None, p, 1, 2, 4, 4, None) == -2                        # 1 or 3 channels
    assert lib.sstem_warp_stitch_u8(p, p, None, p, 1, 3, 3, 3, None) == -2                        # H*W % 4
    assert lib.sstem_sections_to_input(p, p, p, 1, 4, 4, -1, None) == -2
    assert lib.sstem_prediction_to_u8(p, None, 1, 4, 4, 0, None) == -1
    assert lib.sstem_prediction_to_u8(p + 1, p, 1, 4, 4, 0, None) == -3
    assert lib.sstem_frame_mean_pad(None, 48, p, 1, 3, 4, 4, 25, 0, None) == -1
    assert lib.sstem_frame_mean_pad(p, 47, p, 1, 3, 4, 4, 25, 0, None) == -2                     # batch stride < C*H*W
    assert lib.sstem_frame_mean_pad(p, 48, p, 1, 3, 4, 4, 25, 1, None) == -5
    assert lib.sstem_sepconv_forward_tiled(p, p, p, p, 1, 3, 8, 8, 51, 2 | 4, None) == -5        # accumulate + gray replicas
    assert lib.sstem_tap_conv3x3_packed_elems() == 9 * 13 * 64 * 4 + 64 * 4
    assert lib.sstem_tap_conv3x3_pack_weights(None, p, 51, 51, None) == -1
    assert lib.sstem_tap_conv3x3_pack_weights(p, p, 53, 51, None) == -2                           # cin <= 52
    assert lib.sstem_tap_conv3x3(p, p, None, None, 1, 51, 51, 4, 4, 0, None) == -1
    assert lib.sstem_tap_conv3x3(p, p, None, p, 1, 51, 65, 4, 4, 0, None) == -2                   # cout <= 64
    assert lib.sstem_tap_conv3x3(p, p, None, p, 1, 51, 32, 4, 4, 2, None) == -2                   # tiled: 51 taps only
    assert lib.sstem_tap_conv3x3(p, p, None, p, 1, 51, 51, 4, 4, 4, None) == -5
    assert lib.sstem_tap_conv3x3(p, p + 4, None, p, 1, 51, 51, 4, 4, 0, None) == -3


def test_section_8f_host_mirrors_refuse_to_run_without_cuda(built_lib):
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    import sstem_restoration_b200 as pkg
    with pytest.raises(NotImplementedError):            # as SeparableConvolution.py:47-48
        pkg.interpolation_tail(*(torch.zeros((1, 3, 4, 4)) for _ in range(2)), *(torch.zeros((1, 51, 4, 4)) for _ in range(4)))
    with pytest.raises(pkg.SstemError):
        pkg.sff_sim.degradation(np.zeros((256, 256), np.uint8), 256)
    with pytest.raises(pkg.SstemError):
        pkg.sections_to_input(np.zeros((4, 4), np.uint8), np.zeros((4, 4), np.uint8))
    with pytest.raises(pkg.SstemError):
        pkg.sff_sim.gen_flow(8, 8, 1.0, 0.0)
    with pytest.raises(NotImplementedError):
        pkg.tap_conv3x3(torch.zeros((1, 51, 4, 4)), torch.zeros(30208), cin=51, cout=51)
    with pytest.raises(NotImplementedError):
        pkg.pack_tap_conv_weight(torch.zeros((51, 51, 3, 3)))


def test_host_side_fold_line_logic_matches_oracle():
    """get_two_points / gen_line / fold_line_params draw and derive exactly what the oracle restatement does."""
    import math
    import random
    import oracle
    from sstem_restoration_b200 import sff_sim, synth
    for seed in range(5):
        a, b = random.Random(seed), random.Random(seed)
        assert sff_sim.get_two_points(256, 256, 50, 256, a) == oracle.sff_get_two_points(256, 256, 50, 256, b)
        assert a.random() == b.random()                  # same number of draws consumed
    k, bb = sff_sim.gen_line([0, 70], [256, 190])
    assert (k, bb) == synth.gen_line([0, 70], [256, 190])
    prm = sff_sim.fold_line_params(k, bb, 7, 33, 0.02)
    assert prm[2] == math.sqrt(k ** 2 + 1) and prm[6] == math.sin(math.atan(1 / k)) and prm[7] == math.cos(math.atan(1 / k))
    assert sff_sim.fold_line_params(0, 3.0, 7, 33, 0.02)[6] == math.sin(math.atan(1 / 0.000000001))


@pytest.mark.parametrize("seed", [0, 3, 7])
def test_provider_batch_rewind_logic_matches_sequential_calls(monkeypatch, seed):
    """provider_batch draws the batch optimistically and rewinds the generator when a sample is rejected; with the
    kernels replaced by the numpy oracle (no GPU here) its outputs must equal B sequential degradation + noise calls
    bit for bit.  The seeds are chosen so that several samples need 2-5 attempts (data_provider.py:236-241)."""
    import random
    import oracle
    from sstem_restoration_b200 import sff_sim, synth
    crop, offset, det, B = 448, 96, 256, 5
    imgs = np.stack([synth.em_section(crop, crop, 20 + i) for i in range(B)])

    def fake_degrade(t, params, off):
        outs, f2s, stats = [], [], []
        for img, p in zip(t.numpy(), params):
            k, b, _, lw, fw, dk, _, _ = p
            flow, flow2, mask = synth.gen_flow(crop, crop, k, b, int(lw), int(fw), dk, two_flows=True)
            d = (oracle.image_warp_restated(img, flow) * mask).astype(np.uint8)[off:-off, off:-off]
            outs.append(d)
            f2s.append(flow2[off:-off, off:-off])
            stats.append([int((d == 0).sum()), int(d.sum(dtype=np.int64))])
        st = torch.tensor(stats, dtype=torch.int64)
        return torch.from_numpy(np.stack(outs)), torch.from_numpy(np.stack(f2s)), st, st[:, 0].tolist()

    def fake_contrast(t, stats, boxes):
        for img, s, (ran, px, py, bh, bw, *_) in zip(t.numpy(), stats.tolist(), boxes):
            mean = s[1] / float(det * det)
            zero = img == 0
            box = img[int(px):int(px) + int(bh), int(py):int(py) + int(bw)]
            img[int(px):int(px) + int(bh), int(py):int(py) + int(bw)] = ran * (box - mean) + mean
            img[zero] = 0
        return t

    monkeypatch.setattr(sff_sim, "_as_cuda_u8", lambda x: (torch.as_tensor(x).contiguous(), False))
    monkeypatch.setattr(sff_sim, "_launch_degrade_batch", fake_degrade)
    monkeypatch.setattr(sff_sim, "_launch_contrast_batch", fake_contrast)
    rng = random.Random(seed)
    sff, flow2 = sff_sim.provider_batch(torch.from_numpy(imgs), crop, offset, rng=rng)
    ref = random.Random(seed)
    for i in range(B):
        d, f2 = oracle.provider_degradation_restated(imgs[i], crop, offset, ref, 50)
        want = oracle.sff_noise_restated(d, det, ref)
        assert np.array_equal(sff[i].numpy(), want), (seed, i)
        assert np.array_equal(flow2[i].numpy().view(np.uint32), f2.view(np.uint32)), (seed, i)
    assert rng.random() == ref.random()                   # both generators consumed exactly the same draws

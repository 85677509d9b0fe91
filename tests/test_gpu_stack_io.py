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


@pytest.mark.parametrize("C,H,W", [(3, 36, 44), (1, 36, 44), (3, 256, 256)])
def test_warp_stitch_bit_exact(C, H, W):
    """sff_scripts_fusion/inference.py:163-171 (uint8 cast, PIL 'L', stitch mask) on the device."""
    r = np.random.default_rng(C * H)
    w = r.random((2, C, H, W), dtype=np.float32)
    w[:, :, 5:9, :] = 0.004                                    # below 2/255: taken from the interpolated section
    w[0, :, 20, 3] = 1.0
    interp = r.integers(0, 256, (2, H, W), dtype=np.uint8)
    gray, stitch = pkg.warp_stitch(torch.from_numpy(w).cuda(), torch.from_numpy(interp).cuda())
    for b in range(2):
        want_gray, want_stitch = oracle.warp_stitch_restated(w[b], interp[b])
        assert np.array_equal(gray[b].cpu().numpy(), want_gray)
        assert np.array_equal(stitch[b].cpu().numpy(), want_stitch)


def test_sections_to_input_single_section():
    r = np.random.default_rng(9)
    a = r.integers(0, 256, (40, 48), dtype=np.uint8)
    got = pkg.sections_to_input(torch.from_numpy(a).cuda(), None, 3)
    want = oracle.sections_to_input_restated(a, a, 3)[:, :3]
    assert got.shape == (1, 3, 46, 54) and np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_restore_stack_matches_the_op_by_op_expression():
    """restore_stack (uint8 wire, prefetching uploads, fused tail, warp, stitch) == the same steps spelled out with the
    oracle-checked operators one target at a time, bit for bit; also from a CUDA-resident stack and with to_host."""
    N, H, W = 6, 64, 96
    stack = torch.from_numpy(np.stack([synth.em_section(H, W, 20 + i) for i in range(N)]))
    gen = torch.Generator(device="cuda").manual_seed(5)
    taps = [torch.softmax(torch.randn((1, 51, H, W), device="cuda", generator=gen), 1) for _ in range(4)]
    flow_np, _ = synth.random_fold_flow(H, W, 555)
    flow = torch.from_numpy(np.ascontiguousarray(flow_np.transpose(2, 0, 1))[None]).cuda().permute(0, 2, 3, 1)
    seen = []

    def taps_fn(k, x):
        seen.append(k)
        assert x.shape == (1, 6, H, W)
        return taps

    res = pkg.restore_stack(stack.pin_memory(), taps_fn, lambda k, xk, interp: flow, to_host=True)
    assert seen == [1, 2, 3, 4] and res["stats"]["targets"] == 4 and res["stats"]["h2d_bytes"] == N * H * W
    warp = pkg.SpatialTransformation(True)
    for i, k in enumerate(range(1, N - 1)):
        x = pkg.sections_to_input(stack[k - 1].cuda(), stack[k + 1].cuda(), 0)
        interp = pkg.prediction_to_uint8(pkg.interpolation_tail(x[:, :3], x[:, 3:6], *taps), 0)
        warped = warp(pkg.sections_to_input(stack[k].cuda(), None, 0), flow)
        gray, stitch = oracle.warp_stitch_restated(warped[0].cpu().numpy(), interp[0].cpu().numpy())
        assert torch.equal(res["interp"][i], interp[0].cpu())
        assert np.array_equal(res["warped"][i].numpy(), gray) and np.array_equal(res["stitch"][i].numpy(), stitch)
    assert not res["interp"].is_cuda and res["stats"]["d2h_bytes"] == 3 * 4 * H * W
    on_dev = pkg.restore_stack(stack.cuda(), taps_fn, None)
    assert set(on_dev) == {"interp", "stats"} and torch.equal(on_dev["interp"].cpu(), res["interp"])
    # the per-rank shards of a 3-way split (no process group here: each call returns its own shard) tile the result
    parts = [pkg.restore_stack(stack.cuda(), taps_fn, None, rank=r, world_size=3)["interp"] for r in range(3)]
    assert [p.shape[0] for p in parts] == [2, 1, 1] and torch.equal(torch.cat(parts), on_dev["interp"])


def test_restore_stack_with_tile_major_taps_from_the_tap_producer():
    """The N2 chain inside the stack loop: taps_fn returns tile-major taps (written by the tcgen05 tap producer from
    half-resolution activations), restore_stack then takes the tile-major tail -- no [1,51,H,W] tensor anywhere.  Equal,
    up to one uint8 count where fp32 rounding straddles the truncation, to the run on the same taps converted to NCHW."""
    N, H, W = 5, 64, 96
    stack = torch.from_numpy(np.stack([synth.em_section(H, W, 30 + i) for i in range(N)])).cuda()
    gen = torch.Generator(device="cuda").manual_seed(9)
    acts = [torch.relu(torch.randn((1, 51, H // 2, W // 2), device="cuda", generator=gen)) for _ in range(4)]
    prods = [pkg.ModuleTapProducer(tiled=True).cuda() for _ in range(4)]
    plain = [pkg.ModuleTapProducer(tiled=False).cuda() for _ in range(4)]
    for a, b in zip(prods, plain):
        b.load_state_dict(a.state_dict())
    pkg.set_gray_replicated("assert")                   # sections are gray x3 by construction (inference.py:71-74)
    try:
        tiled = pkg.restore_stack(stack, lambda k, x: [m(a) for m, a in zip(prods, acts)], None)["interp"]
        nchw = pkg.restore_stack(stack, lambda k, x: [m(a) for m, a in zip(plain, acts)], None)["interp"]
    finally:
        pkg.set_gray_replicated("off")
    assert tiled.shape == (N - 2, H, W) and tiled.dtype == torch.uint8
    assert (tiled.int() - nchw.int()).abs().max().item() <= 1


def _gather_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from sstem_restoration_b200 import shard
    g = shard.SectionGatherer((3, 8, 8), torch.float32, 2, dev, dst=0)
    ok = True
    for rep in range(3):
        local = torch.stack([torch.full((3, 8, 8), float(10 * rank + j + rep), device=dev) for j in range(2)])
        got = g.gather(local)
        torch.cuda.synchronize()
        if rank == 0:
            ok = ok and [float(got[i, 0, 0, 0]) for i in range(2 * world)] == [float(10 * r + j + rep) for r in range(world) for j in range(2)]
        else:
            ok = ok and got is None
    q.put((rank, ok, g.mode, g.why))
    dist.barrier()
    dist.destroy_process_group()


def test_section_gatherer_peer_writes_on_two_gpus():
    """SectionGatherer on real GPUs: symmetric-memory peer writes (or, where the driver refuses, the NCCL fall-back) give
    the units in rank order on dst, repeatedly."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    print("SectionGatherer mode:", res[0][2], res[0][3])


@pytest.mark.parametrize("C,H,W,kind", [(1, 64, 128, "fold"), (3, 64, 128, "fold"), (3, 40, 64, "noise"), (1, 34, 50, "noise"), (3, 256, 256, "fold")])
def test_warp_and_stitch_fused_equals_the_two_steps(C, H, W, kind):
    """The stitch assembly as the warp kernel's epilogue == SpatialTransformation followed by warp_stitch, bit for bit
    (TMA path, its global-gather tiles on a rough flow, and the scratch-image fall-back for W % 4 != 0)."""
    r = np.random.default_rng(H + W + C)
    moving = torch.from_numpy(r.random((2, C, H, W), dtype=np.float32)).cuda()
    moving[:, :, 10:14, :] = 0.003                                          # < 2/255 after the cast: taken from the interpolation
    if kind == "fold":
        f = synth.random_fold_flow(H, W, 555)[0]
    else:
        f = synth.noise_flow(H, W, 6.0)
    flow = torch.from_numpy(np.ascontiguousarray(f.transpose(2, 0, 1))[None]).cuda().expand(2, 2, H, W).contiguous().permute(0, 2, 3, 1)
    interp = torch.from_numpy(r.integers(0, 256, (2, H, W), dtype=np.uint8)).cuda()
    warped = pkg.SpatialTransformation(True)(moving, flow)
    g0, s0 = pkg.warp_stitch(warped, interp)
    g1, s1 = pkg.warp_and_stitch(moving, flow, interp)
    assert torch.equal(g0, g1) and torch.equal(s0, s1)
    _, s2 = pkg.warp_and_stitch(moving, flow, interp, want_gray=False)
    assert torch.equal(s2, s0)

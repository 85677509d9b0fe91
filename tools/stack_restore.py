"""BASELINE config 5 from the command line: restore a synthetic EM stack, sharded by section pair.

    python tools/stack_restore.py --sections 20 --size 2048            # 1 GPU
    torchrun --nproc-per-node 8 tools/stack_restore.py --sections 100 --size 4096

A thin wrapper over `sstem_restoration_b200.restore_stack` (the package API; `bench.py` measures the same call as
`extra.configs.c5_stack`): target k is interpolated from sections k-1 and k+1 (sff_scripts_interp/inference.py:69-89 ->
model_interp.py:90-97, one fused launch), section k is flow-warped and stitched (sff_scripts_fusion/inference.py:149-171,
one fused launch), uint8 on the wire.  Taps and flow are synthetic: the KPN / flow net are out of scope."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sstem_restoration_b200 as pkg  # noqa: E402
from sstem_restoration_b200 import shard, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sections", type=int, default=20)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--to-host", action="store_true", help="every rank downloads its own restored sections into pinned memory")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    H = W = args.size
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    taps = [torch.softmax(torch.randn((1, 51, H, W), device=dev, generator=gen), 1) for _ in range(4)]
    flow_np, _ = synth.random_fold_flow(H, W, 555)
    flow = torch.from_numpy(np.ascontiguousarray(flow_np.transpose(2, 0, 1))[None]).to(dev).permute(0, 2, 3, 1)
    tile = synth.em_section(min(H, 1024), min(W, 1024), 0)
    base = torch.from_numpy(np.tile(tile, (H // tile.shape[0], W // tile.shape[1])))
    stack = torch.stack([torch.roll(base, shifts=7 * k, dims=1) for k in range(args.sections)]).pin_memory()
    pkg.set_gray_replicated("assert")
    kw = dict(rank=rank, world_size=world, dst=0, device=dev, to_host=args.to_host)
    pkg.restore_stack(stack[:4], lambda k, x: taps, lambda k, xk, interp: flow, device=dev)          # warm-up
    if world > 1:
        shard.gather_sections(torch.zeros((1, 8, 8), dtype=torch.uint8, device=dev), world, dst=0)
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = pkg.restore_stack(stack, lambda k, x: taps, lambda k, xk, interp: flow, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        t = float(ms.item()) * 1e-3
        n = out["stats"]["targets"]
        print(json.dumps({"workload": f"stack restoration, {args.sections} sections {H}x{W}, {n} targets", "n_gpus": world,
                          "seconds": round(t, 4), "sections_per_s": round(n / t, 2), "mpix_per_s": round(n * H * W / t / 1e6, 1),
                          "outputs": {k: list(v.shape) for k, v in out.items() if k != "stats" and v is not None}, "stats": out["stats"]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""torchrun --nproc-per-node N gpurun_scratch/c5_ranks.py: config-5 variants with per-rank times."""
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, "/root/repo")
import bench
import sstem_restoration_b200 as pkg
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
for name, kw in [("nchw", {}), ("host", dict(to_host=True)), ("tiled", dict(tiled_taps=True)), ("tiled", dict(tiled_taps=True))]:
    r = bench.run_c5_stack(pkg, dev, rank, world, dist, 100, 4096, **kw)
    if rank == 0:
        print(name, r["sections_per_s"], r["seconds"], r["per_rank_ms"], flush=True)
dist.destroy_process_group()

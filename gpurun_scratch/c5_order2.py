import sys, os, torch
sys.path.insert(0, "/root/repo")
import bench
import sstem_restoration_b200 as pkg
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
if os.environ.get("NO_EMPTY"):
    torch.cuda.empty_cache = lambda: None
for name, kw in [("nchw", {}), ("nchw", {}), ("host", dict(to_host=True)), ("tiled", dict(tiled_taps=True)), ("nchw", {}), ("tiled", dict(tiled_taps=True))]:
    r = bench.run_c5_stack(pkg, dev, 0, 1, None, 100, 4096, **kw)
    print(name, r["sections_per_s"], r["ms_per_target_on_busiest_rank"], "reserved GB", round(torch.cuda.memory_reserved() / 2**30, 1), flush=True)

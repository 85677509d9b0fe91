"""Times fwd / bwd at a few small shapes (choice of the persistent-kernel threshold).  usage: python tools/sweep_small.py"""
import os, sys, json, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sstem_restoration_b200 as pkg
from sstem_restoration_b200 import _lib
lib = _lib.load()
K = 51
def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (B, H, W) in [(1, 256, 256), (1, 512, 512), (2, 512, 512), (4, 512, 512), (1, 1024, 1024), (8, 512, 512)]:
    C = 3
    inp = torch.rand((B, C, H + 50, W + 50), device="cuda")
    v = torch.softmax(torch.randn((B, K, H, W), device="cuda"), 1); h = torch.softmax(torch.randn((B, K, H, W), device="cuda"), 1)
    g = torch.randn((B, C, H, W), device="cuda"); out = torch.empty((B, C, H, W), device="cuda")
    gv, gh = torch.empty_like(v), torch.empty_like(h)
    st = torch.cuda.current_stream().cuda_stream
    f = timeit(lambda: lib.sstem_sepconv_forward(inp.data_ptr(), v.data_ptr(), h.data_ptr(), out.data_ptr(), B, C, H, W, K, 0, st))
    b = timeit(lambda: lib.sstem_sepconv_backward(g.data_ptr(), inp.data_ptr(), v.data_ptr(), h.data_ptr(), None, gv.data_ptr(), gh.data_ptr(), B, C, H, W, K, 0, st))
    print(json.dumps({"shape": [B, C, H, W], "min_tiles": os.environ.get("SSTEM_V3_MIN_TILES", "default"), "fwd_ms": round(f, 4), "bwd_ms": round(b, 4)}))

"""Runs a few launches of one op for profiling under ncu.  usage: run_one.py <fwd|bwd|gi|warp|tail|tailbwd> B C H W [reps]
(tail / tailbwd: fused interpolation tail, gray-replicated frames asserted when GRAY=1)"""
import os
import sys
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sstem_restoration_b200 as pkg  # noqa: E402
from sstem_restoration_b200 import _lib  # noqa: E402

op = sys.argv[1]
B, C, H, W = (int(a) for a in sys.argv[2:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
K = 51
dev = "cuda"
lib = _lib.load()
torch.manual_seed(0)
st = torch.cuda.current_stream().cuda_stream
if op == "warp":
    im = torch.rand((B, C, H, W), device=dev)
    kind = os.environ.get("FLOW", "noise")
    if kind == "noise":
        fl = (5 * torch.randn((B, 2, H, W), device=dev)).permute(0, 2, 3, 1)
    elif kind == "fold":                                 # the SFF fold flow of the benchmark (gen_flow, seed 555)
        import numpy as np
        from sstem_restoration_b200 import synth
        f = synth.random_fold_flow(H, W, 555)[0]
        fl = torch.from_numpy(np.ascontiguousarray(f.transpose(2, 0, 1))[None]).to(dev).expand(B, 2, H, W).contiguous().permute(0, 2, 3, 1)
    else:
        fl = (torch.zeros((B, 2, H, W), device=dev) + 3.3).permute(0, 2, 3, 1)
    m = pkg.SpatialTransformation(True)
    for _ in range(reps):
        m(im, fl)
elif op == "sff":
    from sstem_restoration_b200 import sff_sim, synth
    img = torch.randint(0, 256, (B, H, W), dtype=torch.uint8, device=dev)
    k, b = synth.gen_line([0, W // 3], [H, 2 * W // 3])
    prm = [sff_sim.fold_line_params(k, b, 12, 60, 0.05)] * B
    for _ in range(reps):
        sff_sim.gen_flow_warp(img, prm, want_flow=False, want_mask=False)
elif op in ("tail", "tailbwd"):
    f1 = torch.rand((B, 1, H, W), device=dev).expand(B, C, H, W).contiguous()
    f2 = torch.rand((B, 1, H, W), device=dev).expand(B, C, H, W).contiguous()
    taps = [torch.softmax(torch.randn((B, K, H, W), device=dev), 1) for _ in range(4)]
    out = torch.empty((B, 1, H, W), device=dev)
    grads = [torch.empty_like(t) for t in taps]
    g = torch.randn((B, 1, H, W), device=dev)
    flag = 2 if os.environ.get("GRAY", "1") == "1" else 0
    for _ in range(reps):
        if op == "tail":
            lib.sstem_interp_tail_forward(f1.data_ptr(), f2.data_ptr(), C * H * W, *[t.data_ptr() for t in taps], out.data_ptr(), B, C, H, W, K, flag, st)
        else:
            lib.sstem_interp_tail_backward(g.data_ptr(), f1.data_ptr(), f2.data_ptr(), C * H * W, *[t.data_ptr() for t in taps],
                                           *[t.data_ptr() for t in grads], B, C, H, W, K, flag, st)
else:
    inp = torch.rand((B, C, H + 50, W + 50), device=dev)
    v = torch.softmax(torch.randn((B, K, H, W), device=dev), 1)
    h = torch.softmax(torch.randn((B, K, H, W), device=dev), 1)
    g = torch.randn((B, C, H, W), device=dev)
    out = torch.empty((B, C, H, W), device=dev)
    gv, gh, gi = torch.empty_like(v), torch.empty_like(h), torch.empty_like(inp)
    for _ in range(reps):
        if op == "fwd":
            lib.sstem_sepconv_forward(inp.data_ptr(), v.data_ptr(), h.data_ptr(), out.data_ptr(), B, C, H, W, K, 0, st)
        elif op == "bwd":
            lib.sstem_sepconv_backward(g.data_ptr(), inp.data_ptr(), v.data_ptr(), h.data_ptr(), None, gv.data_ptr(), gh.data_ptr(), B, C, H, W, K, 0, st)
        elif op == "gi":
            lib.sstem_sepconv_backward(g.data_ptr(), inp.data_ptr(), v.data_ptr(), h.data_ptr(), gi.data_ptr(), None, None, B, C, H, W, K, 0, st)
torch.cuda.synchronize()
print("done", op, B, C, H, W)

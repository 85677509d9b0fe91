"""Warp timing on the benchmark flows (fold / N(0,5px) / constant) at 2048^2 and 4096^2.  usage: python tools/bench_warp.py"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sstem_restoration_b200 as pkg
from sstem_restoration_b200 import synth
st = pkg.SpatialTransformation(True)
dev = "cuda"
for n, nsets in ((2048, 6), (4096, 3)):
    sec = torch.from_numpy(synth.em_section(min(n, 2048), min(n, 2048), 50).astype(np.float32) / 255.0).to(dev)
    if n > 2048: sec = sec.repeat(2, 2)
    for name, flow_np in (("fold", synth.random_fold_flow(n, n, 555)[0]), ("noise5", synth.noise_flow(n, n, 5.0)), ("noise2", synth.noise_flow(n, n, 2.0)),
                          ("const", np.full((n, n, 2), 3.3, np.float32))):
        planar0 = torch.from_numpy(np.ascontiguousarray(flow_np.transpose(2, 0, 1))[None]).to(dev)
        bufs = [(torch.roll(sec, 17 * i, 1)[None, None].expand(1, 3, n, n).contiguous(), (planar0 + 0.01 * i).permute(0, 2, 3, 1)) for i in range(nsets)]
        for _ in range(5):
            for im, fl in bufs: st(im, fl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            for im, fl in bufs: st(im, fl)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (20 * nsets)
        print(json.dumps({"n": n, "flow": name, "ms": round(ms, 5), "gb_per_s": round(32 * n * n / ms / 1e6, 1), "frac_hbm_6538": round(32 * n * n / ms / 1e6 / 6538.3, 4)}))
        del bufs

"""Kernel-level timing of the tap producer (csrc/tapconv.cu) next to the library chain it replaces
(F.interpolate -> cuDNN conv2d in TF32 -> taps_to_tiled), CUDA events, inputs rotated through > L2.

    python tools/bench_tapconv.py [--size 2048] [--reps 10] [--no-lib]

Prints one JSON line per variant: ms, useful TFLOP/s (2 * 51 * 51 * 9 flop per output pixel), fraction of the TF32 peak
(MEASURED_PEAKS.json's dense bf16 figure / 2 when present, else 1100 TFLOP/s) and the HBM-side GB/s (source read + taps written)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sstem_restoration_b200 as pkg  # noqa: E402


def tf32_peak():
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        for k in ("bf16_tflops", "bf16_tflops_sustained"):     # burst figure: these kernels are timed alone
            if k in d:
                return float(d[k]) / 2, "MEASURED_PEAKS.json %s / 2" % k
        flat = json.dumps(d)
        return 1100.0, "nominal (no bf16 key recognised in MEASURED_PEAKS.json: %s)" % flat[:120]
    except Exception:
        return 1100.0, "nominal TF32 dense (B200_PROFILING.md)"


def timeit(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-lib", action="store_true")
    a = ap.parse_args()
    dev = "cuda"
    B, H, W = a.batch, a.size, a.size
    peak, peak_src = tf32_peak()
    torch.manual_seed(0)
    nrot = 4                                               # 4 x (54 MB in + 855 MB out at 2048^2) >> L2
    xs = [torch.relu(torch.randn((B, 51, H // 2, W // 2), device=dev)) for _ in range(nrot)]
    conv = torch.nn.Conv2d(51, 51, 3, 1, 1).to(dev)
    packed = pkg.pack_tap_conv_weight(conv.weight.detach())
    bias = conv.bias.detach()
    flop = 2.0 * 51 * 51 * 9 * B * H * W
    hbm = (B * 51 * (H // 2) * (W // 2) + B * 51 * H * W) * 4
    it = [0]

    def report(name, fn, **extra):
        ms = timeit(fn, a.reps)
        print(json.dumps({"op": name, "out": [B, 51, H, W], "ms": round(ms, 4), "tflops_useful": round(flop / ms / 1e9, 1),
                          "frac_tf32_peak": round(flop / ms / 1e9 / peak, 3), "hbm_GBs": round(hbm / ms / 1e6, 1), **extra}), flush=True)
        return ms

    def nxt():
        it[0] += 1
        return xs[it[0] % nrot]

    print(json.dumps({"tf32_peak_tflops": peak, "source": peak_src}))
    report("tapconv_fused_tiled", lambda: pkg.tap_conv3x3(nxt(), packed, bias, upsample=True, tiled=True))
    report("tapconv_fused_nchw", lambda: pkg.tap_conv3x3(nxt(), packed, bias, upsample=True, tiled=False))
    ups = [torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True) for x in xs[:2]]
    k = [0]

    def nxt_up():
        k[0] += 1
        return ups[k[0] % 2]
    report("tapconv_conv_only_nchw", lambda: pkg.tap_conv3x3(nxt_up(), packed, bias, upsample=False, tiled=False))
    if not a.no_lib:
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cudnn.benchmark = True
        with torch.no_grad():
            report("library_cudnn_conv_tf32_only", lambda: conv(nxt_up()))
            report("library_upsample_conv_tf32", lambda: conv(torch.nn.functional.interpolate(nxt(), scale_factor=2, mode="bilinear", align_corners=True)))
            report("library_upsample_conv_tf32_to_tiled",
                   lambda: pkg.taps_to_tiled(conv(torch.nn.functional.interpolate(nxt(), scale_factor=2, mode="bilinear", align_corners=True))))
            torch.backends.cudnn.allow_tf32 = False
            report("library_upsample_conv_fp32", lambda: conv(torch.nn.functional.interpolate(nxt(), scale_factor=2, mode="bilinear", align_corners=True)))


if __name__ == "__main__":
    main()

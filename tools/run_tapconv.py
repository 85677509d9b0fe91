"""One launch pair of the tap producer at the c4 size, for ncu: python tools/run_tapconv.py [size]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sstem_restoration_b200 as pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
torch.manual_seed(0)
x = torch.relu(torch.randn((1, 51, n // 2, n // 2), device="cuda"))
conv = torch.nn.Conv2d(51, 51, 3, 1, 1).cuda()
packed = pkg.pack_tap_conv_weight(conv.weight.detach())
for _ in range(3):
    out = pkg.tap_conv3x3(x, packed, conv.bias.detach(), upsample=True, tiled=True)
torch.cuda.synchronize()
print(float(out.abs().mean()))

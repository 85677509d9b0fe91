"""compute-sanitizer target: the tap producer on ragged / aligned / multi-tile shapes, both layouts and the plain-conv path."""
import torch
import sstem_restoration_b200 as pkg

torch.manual_seed(0)
w = torch.randn((51, 51, 3, 3), device="cuda") / 21
b = torch.randn(51, device="cuda")
pk = pkg.pack_tap_conv_weight(w)
for (h, ww, ups, tiled) in [(40, 64, True, True), (13, 9, True, False), (37, 29, False, True), (72, 100, True, True)]:
    x = torch.relu(torch.randn((2, 51, h, ww), device="cuda"))
    y = pkg.tap_conv3x3(x, pk, b, upsample=ups, tiled=tiled)
    torch.cuda.synchronize()
    print(h, ww, ups, tiled, float(y.abs().mean()))
w2 = torch.randn((16, 8, 3, 3), device="cuda")
print(float(pkg.tap_conv3x3(torch.randn((1, 8, 20, 12), device="cuda"), pkg.pack_tap_conv_weight(w2), None).abs().mean()))

"""Generates tests/golden/kpn_taps_ref.npz: 51-tap kernels predicted by the REFERENCE's own random-init KPN
(BASELINE config 2 / SURVEY.md section 8d "KPN taps"), run on CPU in the build container.

    python tests/golden/make_kpn_taps_golden.py

`sff_scripts_interp/model/model_interp.py` is imported unmodified with `sstem_restoration_b200/compat` first on
sys.path, so its `from libs.sepconv.SeparableConvolution import SeparableConvolution` (model_interp.py:5) resolves to
the drop-in -- which is also the zero-edit integration path of INTEGRATION.md section 3.  IFNet(kernel_size=51) is built
with torch.manual_seed(0) (orthogonal init, model_interp.py:145-148) and run on a pair of synthetic EM sections; forward
hooks on the four tap branches (upconv51_1..4, model_interp.py:86-89) capture k2h, k2v, k1h, k1v, and the forward then
stops at the sepconv call, which raises NotImplementedError on CPU tensors exactly like the reference's op
(libs/sepconv/SeparableConvolution.py:47-48).  A 32x32 crop of the taps is stored (835 KB); tests rebuild the frames
from the seeds.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SSTEM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sstem_restoration_b200", "compat"))     # libs.sepconv.SeparableConvolution -> drop-in
sys.path.insert(1, os.path.join(REF, "sff_scripts_interp"))

from tests.golden import cases  # noqa: E402


def main():
    from model.model_interp import IFNet                  # the reference's model file, unmodified
    from sstem_restoration_b200 import synth
    p = cases.kpn_taps_case()
    torch.manual_seed(p["torch_seed"])
    net = IFNet(kernel_size=51).eval()
    taps = {}
    for name in ("upconv51_1", "upconv51_2", "upconv51_3", "upconv51_4"):
        getattr(net, name).register_forward_hook(lambda m, i, o, name=name: taps.__setitem__(name, o.detach()))
    x = torch.from_numpy(cases.kpn_frames(p))              # [1,6,256,256]: i1 = x[:, :3], i2 = x[:, 3:6]
    with torch.no_grad():
        try:
            net(x)
            raise SystemExit("expected the CPU sepconv call to raise like the reference's")
        except NotImplementedError:
            pass
    y0, x0, n = p["crop_y"], p["crop_x"], p["crop"]
    crop = lambda t: np.ascontiguousarray(t[:, :, y0:y0 + n, x0:x0 + n].numpy())
    # model_interp.py:86-89: k2h = upconv51_1, k2v = upconv51_2, k1h = upconv51_3, k1v = upconv51_4
    out = {"k2h": crop(taps["upconv51_1"]), "k2v": crop(taps["upconv51_2"]), "k1h": crop(taps["upconv51_3"]), "k1v": crop(taps["upconv51_4"])}
    np.savez_compressed(os.path.join(HERE, "kpn_taps_ref.npz"), **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


if __name__ == "__main__":
    main()

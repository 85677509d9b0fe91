"""Generates tests/golden/tap_producer_ref.npz: input, parameters and output of the last two layers of one of the
REFERENCE's own tap branches (IFNet._kernel_module: ..., upsample, Conv2d(51, 51, 3, 1, 1);
sff_scripts_interp/model/model_interp.py:18, 34, 130-137), run on CPU (true fp32) in the build container.

    python tests/golden/make_tap_producer_golden.py

`model_interp.py` is imported unmodified (see make_kpn_taps_golden.py for the import path trick).  IFNet(kernel_size=51) is
built with torch.manual_seed(3) and run on a 64x64 pair of synthetic sections; hooks on `upconv51_1[6]` (the nn.Upsample
all branches and the decoder share -- its last input before the conv runs is this branch's) and `upconv51_1[7]` (the
final Conv2d) capture what goes in and what comes out.  The forward stops at the
CPU sepconv call, which raises NotImplementedError like the reference's op.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SSTEM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sstem_restoration_b200", "compat"))
sys.path.insert(1, os.path.join(REF, "sff_scripts_interp"))


def main():
    from model.model_interp import IFNet                  # the reference's model file, unmodified
    from sstem_restoration_b200 import synth
    torch.manual_seed(3)
    net = IFNet(kernel_size=51).eval()
    branch = net.upconv51_1
    assert isinstance(branch[6], torch.nn.Upsample) and isinstance(branch[7], torch.nn.Conv2d)
    got = {}
    def hook_conv(m, i, o):                                # a hook that returns a value would replace the output
        got["up"], got["y"] = i[0].detach().clone(), o.detach().clone()

    branch[7].register_forward_hook(hook_conv)
    branch[6].register_forward_hook(lambda m, i, o: got.__setitem__("x", i[0].detach().clone()))   # shared nn.Upsample: last call before
    branch[7].register_forward_pre_hook(lambda m, i: got.__setitem__("x_final", got["x"]))        # the conv is this branch's
    sec = np.stack([synth.em_section(64, 64, index=21 + k) for k in range(2)]).astype(np.float32) / 255.0
    x = torch.from_numpy(np.concatenate([np.repeat(sec[0][None, None], 3, 1), np.repeat(sec[1][None, None], 3, 1)], 1))
    with torch.no_grad():
        try:
            net(x)
            raise SystemExit("expected the CPU sepconv call to raise like the reference's")
        except NotImplementedError:
            pass
    out = {"x": got["x_final"].numpy(), "weight": branch[7].weight.detach().numpy(), "bias": branch[7].bias.detach().numpy(),
           "up_c0_3": got["up"][:, :4].numpy(), "y": got["y"].numpy()}
    np.savez_compressed(os.path.join(HERE, "tap_producer_ref.npz"), **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


if __name__ == "__main__":
    main()

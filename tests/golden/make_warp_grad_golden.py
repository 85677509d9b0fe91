"""Generates tests/golden/warp_torch_grad_ref.npz: gradients of the REFERENCE's own SpatialTransformation
(sff_scripts_unfolding/utils/image_warp_torch.py:5-113, differentiable through its ATen ops) w.r.t. the moving image and
the flow, by autograd on CPU in the build container.

    python tests/golden/make_warp_grad_golden.py

Inputs are the seeded warp cases of tests/golden/cases.py; the upstream gradient is standard normal with
np.random.default_rng(1000 + len(name))."""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SSTEM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(REF, "sff_scripts_unfolding"))

from tests.golden import cases  # noqa: E402


def upstream(name, shape):
    return np.random.default_rng(1000 + len(name)).standard_normal(shape).astype(np.float32)


def main():
    from utils.image_warp_torch import SpatialTransformation      # the reference module, unmodified
    st = SpatialTransformation(False)
    out = {}
    for name, (mv, fl) in cases.warp_torch_cases().items():
        if mv.shape[2] * mv.shape[3] < 4:
            continue
        m = torch.from_numpy(mv).requires_grad_(True)
        f = torch.from_numpy(fl).requires_grad_(True)
        y = st(m, f)
        y.backward(torch.from_numpy(upstream(name, tuple(y.shape))))
        out[name + "_gm"] = m.grad.numpy()
        out[name + "_gf"] = f.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "warp_torch_grad_ref.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

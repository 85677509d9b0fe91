"""Generates tests/golden/sepconv_ref.npz by running the REFERENCE's own CUDA kernels
(libs/sepconv/src/SeparableConvolution_kernel.cu compiled verbatim for sm_100a into
oracle/_ref/libref_sepconv.so) on a B200:

    gpurun -- 'python tests/golden/make_sepconv_ref_golden.py gpurun_out/sepconv_ref.npz'

then copy gpurun_out/sepconv_ref.npz to tests/golden/.  Inputs are rebuilt from the
seeds in tests/golden/cases.py, so the fixture holds reference OUTPUTS only.  The
reference backward is only defined for C == 3 (kernel.cu:100-108)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from tests import _ref_cuda  # noqa: E402
from tests.golden import cases  # noqa: E402


def main(dst):
    out = {}
    for name, p in cases.sepconv_cases().items():
        if p["C"] != 3:
            continue
        inp, v, h, g = (torch.from_numpy(a).cuda() for a in cases.sepconv_inputs(**p))
        o = _ref_cuda.forward(inp, v, h)
        gi, gv, gh = _ref_cuda.backward(g, inp, v, h)
        out[name + "_out"] = o.cpu().numpy()
        out[name + "_gv"] = gv.cpu().numpy()
        out[name + "_gh"] = gh.cpu().numpy()
        out[name + "_gi"] = gi.cpu().numpy()
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "sepconv_ref.npz"))

"""Generates tests/golden/warp_*.npz and flow_*.npz by running the REFERENCE's own
Python (imported from /root/reference; only possible in the build container).

    python tests/golden/make_golden.py

Fixtures hold the seeded inputs' parameters and the reference outputs; tests
rebuild the inputs from the seeds, run the oracle restatement (CPU) and the CUDA
kernels (GPU) and require bit-equality.  Reference functions exercised:
  * simu_sff/image_warp.py:3-111                         image_warp
  * sff_scripts_unfolding/utils/image_warp_torch.py:5-113 SpatialTransformation
  * simu_sff/flow_synthesis.py:13-83                      gen_line, gen_flow
    (imported with a stub `matplotlib` because flow_synthesis.py:6 imports pyplot)
  * simu_sff/simuSFF.py:96-144                            degradation, noise
    (imported with stub `matplotlib` / `skimage` modules: simuSFF.py:10-11 and flow_display.py:2
    import them for PNG I/O and flow colouring only)
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SSTEM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from tests.golden import cases  # noqa: E402


def _load(path, name, extra_path=None):
    if extra_path:
        sys.path.insert(0, extra_path)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if extra_path:
        sys.path.remove(extra_path)
    return mod


def main():
    ref_np = _load(os.path.join(REF, "simu_sff", "image_warp.py"), "ref_image_warp")
    ref_t = _load(os.path.join(REF, "sff_scripts_unfolding", "utils", "image_warp_torch.py"), "ref_image_warp_torch")
    # flow_synthesis imports matplotlib.pyplot and PIL at module level; stub what is absent
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    ref_fs = _load(os.path.join(REF, "simu_sff", "flow_synthesis.py"), "ref_flow_synthesis",
                   extra_path=os.path.join(REF, "simu_sff"))

    # ---- numpy image_warp ----------------------------------------------------------
    out = {}
    for name, (im, flow, mode) in cases.image_warp_cases().items():
        out[name] = ref_np.image_warp(im, flow, mode)
    np.savez_compressed(os.path.join(HERE, "warp_numpy_ref.npz"), **out)
    print("image_warp cases:", {k: v.shape for k, v in out.items()})

    # ---- torch SpatialTransformation (CPU run of the reference) ----------------------
    out = {}
    st = ref_t.SpatialTransformation(use_gpu=False)
    for name, (moving, flow) in cases.warp_torch_cases().items():
        with torch.no_grad():
            o = st(torch.from_numpy(moving), torch.from_numpy(flow))
        out[name] = o.contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "warp_torch_ref.npz"), **out)
    print("SpatialTransformation cases:", {k: v.shape for k, v in out.items()})

    # ---- gen_line / gen_flow ----------------------------------------------------------
    out = {}
    for name, (h, w, p1, p2, lw, fw, dk) in cases.gen_flow_cases().items():
        k, b = ref_fs.gen_line(p1, p2)
        flow, mask = ref_fs.gen_flow(h, w, k, b, lw, fw, dk)
        out[name + "_kb"] = np.array([k, b], np.float64)
        out[name + "_flow"] = flow
        out[name + "_mask"] = mask.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "gen_flow_ref.npz"), **out)
    print("gen_flow cases:", [k for k in out if k.endswith("_flow")])

    # ---- simuSFF.degradation + noise (the reference's CPU-runnable config 1) -------------
    import random
    for name in ("skimage", "skimage.io"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["skimage"], "io"):
        sys.modules["skimage"].io = sys.modules["skimage.io"]
    ref_sim = _load(os.path.join(REF, "simu_sff", "simuSFF.py"), "ref_simuSFF", extra_path=os.path.join(REF, "simu_sff"))
    from sstem_restoration_b200 import synth
    out = {}
    for name, (size, index, seed) in cases.simu_sff_cases().items():
        img = synth.em_section(size, size, index)
        random.seed(seed)
        deformed, flow, mask = ref_sim.degradation(img, size)
        out[name + "_deformed"] = deformed
        # the float32 flow is 0.5-0.7 MB per case: the fixture keeps its SHA-256 (bit-equality is all that is tested)
        out[name + "_flow_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(flow).tobytes()).digest(), np.uint8)
        out[name + "_mask"] = np.packbits(mask.astype(np.uint8))
        out[name + "_noise"] = ref_sim.noise(deformed.copy(), size)
    # ---- the training data providers' degradation + noise: the reference's own method source, executed ----
    # data_provider.py cannot be imported (cv2, h5py, tifffile, torchvision...); the two methods are plain numpy +
    # random, so their source text is cut out of the file and exec'd unmodified with the reference's gen_line /
    # gen_flow (utils/flow_synthesis.py, the three-output variant) and image_warp in scope.
    import textwrap
    for name, (crop, offset, index, seed, which) in cases.provider_degradation_cases().items():
        root = os.path.join(REF, "sff_scripts_" + which)
        fs = _load(os.path.join(root, "utils", "flow_synthesis.py"), "ref_fs_" + which)
        src = open(os.path.join(root, "data", "data_provider.py")).read()
        a = src.index("\tdef degradation(self, img):")
        b = src.index("\t@staticmethod", a)
        ns = {"np": np, "random": random, "gen_line": fs.gen_line, "gen_flow": fs.gen_flow, "image_warp": ref_np.image_warp}
        exec(textwrap.dedent(src[a:b].replace("\t", "    ")), ns)
        me = types.SimpleNamespace(crop_size=[crop, crop], offset=offset, det_size=crop - 2 * offset)
        img = synth.em_section(crop, crop, index)
        random.seed(seed)
        deformed, flow2 = ns["degradation"](me, img)
        out[name + "_deformed"] = deformed
        out[name + "_flow2_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(flow2).tobytes()).digest(), np.uint8)
        out[name + "_noise"] = ns["noise"](me, deformed.copy())
    # the three-output gen_flow on its own (small, kept whole)
    fs = _load(os.path.join(REF, "sff_scripts_unfolding", "utils", "flow_synthesis.py"), "ref_fs_unfolding2")
    k, b = fs.gen_line([0, 20], [64, 60])
    f1, f2, m = fs.gen_flow(64, 80, k, b, 5, 30, 0.05)
    out["gen_flow3_flow"], out["gen_flow3_flow2"], out["gen_flow3_mask"] = f1, f2, m.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "simu_sff_ref.npz"), **out)
    print("simuSFF cases:", {k: v.shape for k, v in out.items() if k.endswith("_noise")})


if __name__ == "__main__":
    main()

"""Builds libsstem_b200.so in-tree with nvcc for sm_100a (no torch headers, no JIT cache).

The library is the whole product: there is no CPU or PyTorch fallback.  Replaces the
reference's `libs/sepconv/install.bash:14-18` (nvcc for compute_37) and
`install.py:23-37` (torch.utils.ffi.create_extension).
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libsstem_b200.so")
SOURCES = ["abi.cu", "sepconv_generic.cu", "sepconv_k51.cu", "warp.cu", "probe.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "--use_fast_math=false", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsstem_b200.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "sstem_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into one shared library; returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + ["-o", LIB_PATH + ".tmp"]
    cmd += os.environ.get("SSTEM_NVCC_EXTRA", "").split()        # experiments only, e.g. -DSSTEM_BWD_ROWS=6
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(log)
    return LIB_PATH

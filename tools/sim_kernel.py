"""Compile sepconv_k51.cu with extra flags and simulate the steady-state loop of a kernel with 1-3 warps/SMSP.
usage: python tools/sim_kernel.py <kernel-regex> <n_ffma2_in_loop> [nvcc flags...]"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import sass_sim as S  # noqa: E402
from sass_stalls import load, ctrl  # noqa: E402


def main():
    pat, nf = sys.argv[1], int(sys.argv[2])
    flags = sys.argv[3:]
    src = os.path.join(HERE, "..", "sstem_restoration_b200", "csrc", "sepconv_k51.cu")
    out = "/tmp/sim_k51.cubin"
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-cubin", "-o", out, src] + flags
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        print(r.stderr[-3000:])
        sys.exit(1)
    ins = load(out, pat)
    S.LAT["LDG"] = 30
    best = None
    for a, t, w0, w1 in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            body = [x for x in ins if lo <= x[0] <= a]
            if sum("FFMA2" in x[1] for x in body) == nf:
                best = body
    if best is None:
        print("steady loop not found")
        sys.exit(1)
    st = sum(max(ctrl(x[3])[0], 1) for x in best)
    print(f"{len(best)} instr, static stall sum {st}, FFMA2 {nf}")
    for nw in (1, 2, 3):
        c = S.simulate(best, nw)
        print(f"  {nw} warps: {c/nw:.0f} cycles per warp-iteration, pipe util {2*nf*nw/c*100:.0f}%")


if __name__ == "__main__":
    main()

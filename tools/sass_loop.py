"""Prints the SASS of the largest loop (by FFMA2 count) of a kernel in an object / shared library, plus its mix.
usage: python tools/sass_loop.py file.o|lib.so <kernel-regex> [--dump]"""
import re
import subprocess
import sys
from collections import Counter


def kernel_sass(path, pat):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = re.search(pat, line) is not None
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(t):
    p = t.split()
    if p[0].startswith("@"):
        p = p[1:]
    op = p[0].split(".")
    if op[0] in ("LDS", "LDG", "STG", "STS") and len(op) > 1 and op[-1] in ("64", "128"):
        return op[0] + "." + op[-1]
    return op[0]


def main():
    path, pat = sys.argv[1], sys.argv[2]
    ins = kernel_sass(path, pat)
    best = None
    for addr, text in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            lo = int(m.group(1), 16)
            body = [(a, t) for a, t in ins if lo <= a <= addr]
            n = sum(1 for _, t in body if "FFMA2" in t)
            # the innermost hot loop: the smallest loop that still holds a step's worth of packed FMAs
            if n >= 40 and (best is None or len(body) < len(best[1])):
                best = (n, body)
    if best is None:
        print("no loop found")
        return
    n, body = best
    c = Counter(opcode(t) for _, t in body)
    reuse = sum(t.count(".reuse") for _, t in body if "FFMA2" in t)
    nonf = len(body) - n
    print(f"{len(ins)} instructions in kernel; main loop {len(body)} instr, FFMA2 {n}, other {nonf}, .reuse flags on FFMA2 {reuse}")
    print(dict(c.most_common(20)))
    if "--dump" in sys.argv:
        for a, t in body:
            print(f"{a:05x} {t}")


if __name__ == "__main__":
    main()

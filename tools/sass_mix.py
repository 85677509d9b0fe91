"""Instruction mix of one kernel (regex on the mangled name) from `cuobjdump -sass`, optionally per line range.
usage: python tools/sass_mix.py lib.so <kernel-regex> [--loops]"""
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


def opcode(ins):
    parts = ins.split()
    if parts[0].startswith("@"):
        parts = parts[1:]
    return parts[0].split(".")[0] if parts[0] not in ("LDS", "LDG", "STG", "STS") else parts[0]


def main():
    path, pat = sys.argv[1], sys.argv[2]
    ins = kernel_sass(path, pat)
    print(f"{len(ins)} instructions")
    if "--loops" in sys.argv:
        # backward branches delimit loops
        for addr, text in ins:
            m = re.search(r"BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                lo = int(m.group(1), 16)
                body = [t for a, t in ins if lo <= a <= addr]
                c = Counter(opcode(t) for t in body)
                print(f"loop {lo:#x}..{addr:#x}: {len(body)} instr:", dict(c.most_common(14)))
    else:
        c = Counter(opcode(t) for _, t in ins)
        print(dict(c.most_common(30)))


if __name__ == "__main__":
    main()

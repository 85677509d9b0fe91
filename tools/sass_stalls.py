"""Decode the control words (stall count, yield, barriers) of one kernel's SASS and estimate the
single-warp issue time of its loops (sum of stall counts; variable-latency waits ignored).
usage: python tools/sass_stalls.py lib.so <kernel-regex> [--dump lo hi]"""
import re
import subprocess
import sys


def load(path, pat):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out, on, cur = [], False, None
    for line in txt.splitlines():
        if "Function :" in line:
            on = re.search(pat, line) is not None
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", line)
        if m:
            cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
            out.append(cur)
            continue
        m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", line)
        if m and cur is not None and cur[3] is None:
            cur[3] = int(m.group(1), 16)
    return [c for c in out if c[3] is not None]


def ctrl(hi):
    stall = (hi >> 41) & 0xF
    yld = (hi >> 45) & 1
    wbar = (hi >> 46) & 7
    rbar = (hi >> 49) & 7
    wait = (hi >> 52) & 0x3F
    return stall, yld, wbar, rbar, wait


def main():
    ins = load(sys.argv[1], sys.argv[2])
    print(len(ins), "instructions")
    if "--dump" in sys.argv:
        i = sys.argv.index("--dump")
        lo, hi = int(sys.argv[i + 1], 16), int(sys.argv[i + 2], 16)
        for a, t, w0, w1 in ins:
            if lo <= a <= hi:
                s, y, wb, rb, wt = ctrl(w1)
                print(f"{a:#06x} st={s:2d} y={y} wb={wb} rb={rb} wait={wt:06b}  {t[:90]}")
        return
    for a, t, w0, w1 in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            body = [(x, ctrl(x[3])) for x in ins if lo <= x[0] <= a]
            n = len(body)
            stalls = sum(c[0] for _, c in body)
            nf2 = sum("FFMA2" in x[1] for x, _ in body)
            f2st = sum(c[0] for x, c in body if "FFMA2" in x[1])
            ldsst = sum(c[0] for x, c in body if re.match(r"(@\S+\s+)?LDS", x[1]))
            nlds = sum(1 for x, c in body if re.match(r"(@\S+\s+)?LDS", x[1]))
            print(f"loop {lo:#x}..{a:#x}: {n} instr, sum(stall)={stalls}, FFMA2 {nf2} (stall sum {f2st}), LDS {nlds} (stall sum {ldsst})")


if __name__ == "__main__":
    main()

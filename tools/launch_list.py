"""Markdown launch list from `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <cmd>`.
usage: python tools/launch_list.py launches.csv "title" [out.md]"""
import csv
import sys
from collections import defaultdict


def main():
    path, title = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else None
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    col = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr or len(r) < len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = agg[r[col["Kernel Name"]][:90]]
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    lines = [f"# {title}", "", "`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and serialised; "
             "compare SHARES, not absolutes.", "", "| kernel | launches | total ms | share | avg ms |", "|---|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% | {ms / n:.4f} |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    else:
        print(text)


if __name__ == "__main__":
    main()

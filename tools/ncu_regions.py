"""Aggregates an `ncu --page source --csv` dump by code region (regions split at backward-branch loop bounds).
usage: ncu -i rep --page source --csv > src.csv ; python tools/ncu_regions.py src.csv"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
data = rows[hdr_i + 1:]
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
base = int(data[0][0], 16)
ins = []
for r in data:
    if len(r) < len(hdr):
        continue
    ins.append(dict(addr=int(r[0], 16) - base, src=r[col["Source"]].strip(), samples=int(r[col["# Samples"]] or 0),
                    execd=int(r[col["Instructions Executed"]] or 0), stalls={n: int(r[col[n]] or 0) for n in stall_cols}))
# loops: backward branches
bounds = set([0])
for i in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", i["src"])
    if m:
        tgt = int(m.group(1), 16)
        tgt = tgt - base if tgt >= base else tgt
        if tgt < i["addr"]:
            bounds.add(tgt)
            bounds.add(i["addr"] + 16)
bounds = sorted(bounds) + [10 ** 9]
total = sum(i["samples"] for i in ins)
print(f"total samples {total}, instructions {len(ins)}")
for lo, hi in zip(bounds, bounds[1:]):
    seg = [i for i in ins if lo <= i["addr"] < hi]
    if not seg:
        continue
    s = sum(i["samples"] for i in seg)
    if s < 0.003 * total:
        continue
    ex = sum(i["execd"] for i in seg)
    agg = {}
    for i in seg:
        for k, v in i["stalls"].items():
            agg[k] = agg.get(k, 0) + v
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:6]
    ops = {}
    for i in seg:
        op = i["src"].split()[1] if i["src"].startswith("@") else i["src"].split()[0]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + i["execd"]
    topops = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    print(f"[{lo:#06x},{hi if hi < 10**9 else 0:#06x}) n={len(seg):4d} samples={s:7d} ({100*s/total:5.1f}%) execd={ex:.3e} "
          + " ".join(f"{k[6:]}={v}" for k, v in top if v) + " | " + " ".join(f"{k}:{v:.2e}" for k, v in topops))
if "--top" in sys.argv:
    for i in sorted(ins, key=lambda i: -i["samples"])[:25]:
        print(f"{i['addr']:#06x} {i['samples']:6d} {i['src'][:90]}  " + " ".join(f"{k[6:]}={v}" for k, v in sorted(i['stalls'].items(), key=lambda kv: -kv[1])[:3] if v))

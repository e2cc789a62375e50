#!/usr/bin/env python
"""Fold an `ncu --page source --csv` dump of one kernel into address regions: warp instructions executed, FP64-pipe
share, stall samples and the top stall reasons per region.
Usage: ncu_regions.py source.csv name=lo-hi [name=lo-hi ...]   (hex addresses, inclusive; everything else = "other")
The addresses are those of `cuobjdump -sass` / tools/sass_blocks.py on the same build."""
import collections
import csv
import sys

path = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    name, rng = a.split("=")
    lo, hi = rng.split("-")
    regions.append((name, int(lo, 16), int(hi, 16)))
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")
stall_cols = [k for k in col if k.startswith("stall_") and "Not" not in k and "not_issued" not in k.lower()]
acc = collections.OrderedDict((n, dict(ex=0, fp=0, s=0, st=collections.Counter())) for n, _, _ in regions)
acc["other"] = dict(ex=0, fp=0, s=0, st=collections.Counter())
base = None
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[col["Address"]], 16)
    if base is None:
        base = addr
    off = addr - base
    name = next((n for n, lo, hi in regions if lo <= off <= hi), "other")
    toks = r[col["Source"]].split()
    op = toks[0] if toks and not toks[0].startswith("@") else (toks[1] if len(toks) > 1 else "?")
    e = int(r[col["Instructions Executed"]] or 0)
    s = int(r[col["# Samples"]] or 0)
    a = acc[name]
    a["ex"] += e
    a["s"] += s
    if op.split(".")[0] in FP64:
        a["fp"] += e
    for k in stall_cols:
        v = int(r[col[k]] or 0)
        if v:
            a["st"][k[6:]] += v
tot_e = sum(a["ex"] for a in acc.values())
tot_s = sum(a["s"] for a in acc.values())
tot_f = sum(a["fp"] for a in acc.values())
print("| region | warp-instructions executed | share | FP64-pipe share | samples | share of time | top stall reasons |")
print("|---|---|---|---|---|---|---|")
for n, a in acc.items():
    if not a["ex"]:
        continue
    ss = sum(a["st"].values()) or 1
    top = ", ".join(f"{k} {100 * v / ss:.0f}%" for k, v in a["st"].most_common(5))
    print(f"| {n} | {a['ex']:,} | {100 * a['ex'] / tot_e:.1f}% | {100 * a['fp'] / a['ex']:.0f}% | {a['s']:,} | "
          f"{100 * a['s'] / tot_s:.1f}% | {top} |")
print(f"\nTotals: {tot_e:,} warp-instructions, {tot_f:,} on the FP64 pipe ({100 * tot_f / tot_e:.1f}%)")

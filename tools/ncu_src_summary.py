#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump of one kernel: opcode mix weighted by executed count,
stall samples by opcode, and the hottest SASS lines.  Usage: ncu_src_summary.py file.csv [topN]"""
import csv
import collections
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = rows[hdr_i + 1:]
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX", "F2F.F64", "MUFU.RCP64H", "MUFU.RSQ64H", "I2F.F64", "F2I.F64",
        "D2I", "I2D")
ex = collections.Counter()
smp = collections.Counter()
tot_ex = tot_s = 0
lines = []
for r in data:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    toks = src.split()
    op = toks[0] if toks and not toks[0].startswith("@") else (toks[1] if len(toks) > 1 else "?")
    op0 = op.split(".")[0]
    e = int(r[col["Instructions Executed"]] or 0)
    s = int(r[col["# Samples"]] or 0)
    ex[op0] += e
    smp[op0] += s
    tot_ex += e
    tot_s += s
    lines.append((s, e, r[col["Address"]][-5:], src, {k[6:]: int(r[col[k]] or 0) for k in col if k.startswith("stall_") and "Not" not in k}))
print(f"total warp instructions {tot_ex:,}  samples {tot_s:,}")
print(f"{'opcode':12s} {'executed':>14s} {'share':>7s} {'samples':>9s} {'share':>7s}")
for op, e in ex.most_common(30):
    print(f"{op:12s} {e:14,d} {100*e/tot_ex:6.2f}% {smp[op]:9,d} {100*smp[op]/max(tot_s,1):6.2f}%")
d = sum(e for op, e in ex.items() if op in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
print(f"FP64-pipe (DFMA/DADD/DMUL/DSETP/DMNMX) share of issued: {100*d/tot_ex:.1f}%   "
      f"issue-cycle model 2*fp64+other: fp64 busy <= {100*2*d/(2*d+tot_ex-d):.1f}%")
print("\nhottest lines")
for s, e, a, src, st in sorted(lines, key=lambda x: -x[0])[:top]:
    t = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{s:7d} {e:12,d} {a} {src[:70]:70s} {t}")

#!/usr/bin/env python
"""Per-basic-block opcode mix of one kernel in an object file (cuobjdump -sass).
Usage: sass_blocks.py obj kernel_substring [min_block_len]"""
import collections
import re
import subprocess
import sys

obj, key = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 16
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
on = False
ins = []
for l in txt.splitlines():
    if "Function :" in l:
        on = key in l
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
targets = set()
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)*0x([0-9a-f]+)", t)
    if m:
        targets.add(int(m.group(1), 16))


def opc(t):
    t = t.split()
    op = t[1] if t[0].startswith("@") else t[0]
    return op.split(".")[0]


FP = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")
blk = collections.Counter()
start = None
print(f"{len(ins)} instructions")
for i, (a, t) in enumerate(ins):
    if start is None:
        start = a
    o = opc(t)
    blk[o] += 1
    nxt = ins[i + 1][0] if i + 1 < len(ins) else None
    if o in ("BRA", "EXIT", "BAR", "RET", "BSYNC") or nxt in targets:
        n = sum(blk.values())
        fp = sum(v for k, v in blk.items() if k in FP)
        if n >= minlen:
            print(f"{start:#07x}-{a:#07x} n={n:4d} fp64={fp:4d} other={n - fp:4d} :", dict(blk.most_common(11)), "|", t[:40])
        blk = collections.Counter()
        start = None

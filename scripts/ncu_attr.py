#!/usr/bin/env python
"""Attribute ncu 'Instructions Executed' (SASS page CSV) to source lines/regions via nvdisasm -g.

usage: ncu_attr.py <sass.csv from `ncu -i rep --page source --csv --print-source sass`> <dis.txt from
`nvdisasm -g -c cubin`> <kernel mangled-name substring> <warps> [region spec file:lo-hi=name ...]
"""
import collections
import csv
import re
import sys

sass, dis, kname, warps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
regions = []
for spec in sys.argv[5:]:
    loc, name = spec.split("=")
    f, rng = loc.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
lines = open(dis).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l][0]
cur, a2l = None, {}
for l in lines[start + 1:]:
    if l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        a2l[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(sass)))
h = rows[1]
ai, ii, si = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
base = None
byline, byreg, byop, samp = (collections.Counter() for _ in range(4))
for r in rows[2:]:
    if len(r) <= ii or not r[ai].startswith("0x"):
        continue
    a = int(r[ai], 16)
    base = a if base is None else base
    n, s = int(r[ii] or 0), int(r[si] or 0)
    cur, op = a2l.get(a - base, (None, "?"))
    f, l = cur if cur else ("?", 0)
    byline[(f, l)] += n
    samp[(f, l)] += s
    name = f
    for rf, lo, hi, rn in regions:
        if rf in f and lo <= l <= hi:
            name = rn
    byreg[name] += n
    t = op.split()
    byop[t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")] += n
tot = sum(byline.values())
print(f"total {tot} = {tot / warps:.1f} per warp")
print("-- regions")
for k, v in byreg.most_common():
    print(f"{v / warps:8.1f} {k}")
print("-- opcodes")
for k, v in byop.most_common(20):
    print(f"{v / warps:8.1f} {k}")
print("-- lines")
for (f, l), v in byline.most_common(40):
    print(f"{v / warps:8.1f} {samp[(f, l)]:6d} {f}:{l}")

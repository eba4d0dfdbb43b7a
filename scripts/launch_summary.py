#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, gzip, io, sys
path = sys.argv[1]
f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r; ki = r.index("Kernel Name"); vi = r.index("Metric Value"); ui = r.index("Metric Unit")
        continue
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    agg[r[ki][:80]][0] += 1; agg[r[ki][:80]][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{v[1]:10.1f}us {v[0]:5d} {v[1]/v[0]:8.2f} {100*v[1]/tot:5.1f}% {k}")

#!/usr/bin/env python3
"""Aggregate an ncu launch list (csv, gpu__time_duration.sum) per kernel over the LAST step of tools/one_step.py."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]; ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
recs = []
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    m = re.search(r"([A-Za-z_0-9]+)\s*(?:<[^(]*>)?\(", r[ki].replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", ""))
    name = m.group(1) if m else r[ki][:40]
    recs.append((name, float(r[vi].replace(",", "")) / 1e3))
half = recs[len(recs) // 2:] if len(sys.argv) < 3 else recs
agg = collections.OrderedDict()
for n, t in half:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(half)} launches, {tot / 1e3:.3f} ms (ncu per-launch times: cold cache, serialised)")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:34s} {c:4d} launches {t:9.1f} us {t / tot * 100:5.1f}%")

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [min_ID] > profiles/<name>.md"""
import csv, re, sys, collections
path = sys.argv[1]; min_id = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if int(r["ID"]) >= min_id:
        rows.append((r["Kernel Name"], r["Grid Size"], r["Block Size"], float(r["Metric Value"])))
def short(n):
    n = re.sub(r"^void ", "", n); n = re.sub(r"\(.*$", "", n)
    n = n.replace("tfq::<unnamed>::", "").replace("<unnamed>::", "")
    return n[:110]
agg = collections.OrderedDict()
for n, g, b, t in rows:
    k = (short(n), g, b)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"launches: {len(rows)}  total device time: {tot*1e-6:.3f} ms (ncu per-launch, cold-cache, serialised: compare SHARES)\n")
print("| kernel | grid | block | launches | total ms | avg us | share |\n|---|---|---|---|---|---|---|")
for (n, g, b), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {g} | {b} | {c} | {t*1e-6:.3f} | {t/c*1e-3:.1f} | {100*t/tot:.1f}% |")

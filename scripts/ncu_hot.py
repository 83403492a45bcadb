#!/usr/bin/env python
"""Top stalled SASS instructions of an ncu report: python scripts/ncu_hot.py rep.ncu-rep [N]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(lines[start:]))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
tot = sum(int(r["# Samples"] or 0) for r in rows)
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
print("total samples", tot)
for idx, r in enumerate(rows): r["_i"] = idx
for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:N]:
    st = sorted(((int(r[c] or 0), c) for c in stall_cols), reverse=True)[:3]
    print(f'{r["_i"]:5d} {100*int(r["# Samples"])/tot:5.1f}%  {r["Source"].strip()[:70]:70s} ' + " ".join(f"{c[6:]}={v}" for v, c in st if v))

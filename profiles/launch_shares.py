#!/usr/bin/env python
"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: python profiles/launch_shares.py profiles/rNN_launches.csv [--each KERNEL_SUBSTRING]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
agg = collections.defaultdict(lambda: [0, 0.0])
each = sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--each" else None
per = []
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    d = dict(zip(H, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = d["Kernel Name"].split("(")[0]
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1e3 if u.startswith("us") else (v / 1e6 if u.startswith("ns") else v)
    agg[k][0] += 1
    agg[k][1] += v
    if each and each in k:
        per.append((d["Grid Size"], v))
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:45s} launches {n:4d}  total {t:10.3f} ms  share {100 * t / tot:5.1f}%")
if per:
    print("per launch:", " ".join(f"{v:.2f}" for _, v in per[:64]))

#!/usr/bin/env python3
"""Condense an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel count / total / share.
usage: tools/ncu_launches.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"'))]
agg = collections.OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    us = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else (v if r["Metric Unit"] in ("us", "usecond") else v * 1e3)
    key = (r["Kernel Name"].split("(")[0][:100], r["Grid Size"], r["Block Size"])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches captured, {tot:.1f} us total (cold-cache, serialised: shares matter, not absolutes)")
print(f"{'kernel':100s} {'grid':>16s} {'block':>14s} {'n':>5s} {'avg us':>10s} {'share':>7s}")
for (k, g, b), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:100s} {g:>16s} {b:>14s} {n:5d} {us / n:10.2f} {100 * us / tot:6.1f}%")

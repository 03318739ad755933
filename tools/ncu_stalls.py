#!/usr/bin/env python3
"""Aggregate the per-instruction warp-stall samples of an .ncu-rep by SASS opcode and by stall reason.
usage: tools/ncu_stalls.py file.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
print(lines[0][:160])
rd = csv.DictReader(io.StringIO("\n".join(lines[1:])))
by_op = defaultdict(lambda: defaultdict(float))
reasons = defaultdict(float)
total = 0.0
shared = defaultdict(lambda: [0.0, 0.0, 0.0])
inst = defaultdict(float)
for r in rd:
    src = r["Source"].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    op = op.rstrip(";")
    n = float(r["# Samples"] or 0)
    total += n
    inst[op] += float(r["Instructions Executed"] or 0)
    for k, v in r.items():
        if k.startswith("stall_") and "Not Issued" not in k and v:
            by_op[op][k] += float(v)
            reasons[k] += float(v)
    if r.get("L1 Wavefronts Shared"):
        s = shared[op]
        s[0] += float(r["L1 Wavefronts Shared"] or 0)
        s[1] += float(r["L1 Wavefronts Shared Ideal"] or 0)
        s[2] += float(r["L1 Wavefronts Shared Excessive"] or 0)
print(f"total samples {total:.0f}")
print("by reason:", ", ".join(f"{k[6:]}={v / total:.1%}" for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:10]))
print("by opcode (share of samples, instructions executed, top reasons):")
for op, d in sorted(by_op.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
    s = sum(d.values())
    tops = ", ".join(f"{k[6:]}={v / s:.0%}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:4])
    print(f"  {op:22s} {s / total:6.1%}  inst={inst[op]:12.0f}  {tops}")
print("shared-memory wavefronts by opcode (actual / ideal / excessive):")
for op, s in sorted(shared.items(), key=lambda kv: -kv[1][0])[:8]:
    if s[0]:
        print(f"  {op:22s} {s[0]:12.0f} {s[1]:12.0f} {s[2]:12.0f}")

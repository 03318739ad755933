#!/usr/bin/env python3
"""Read tuner CSVs, print the best configuration per case and back-end, and write tools/rows_overrides.json.
usage: tools/tune/pick_best.py gpurun_out/tune_*.csv [--write]"""
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
rows = []
for p in sys.argv[1:]:
    if p.startswith("--"):
        continue
    for r in csv.DictReader(l for l in open(p) if not l.startswith("#")):
        if r.get("ok") == "ok":
            rows.append(r)
best = {}
for r in rows:
    key = (int(r["dim"]), r["dtype"], int(r["nq"]), r["backend"])
    if key not in best or float(r["ms_min"]) < float(best[key]["ms_min"]):
        best[key] = r
over = {}
print(f"{'case':12s} {'rows best':34s} {'pipe best':34s}")
for dim in (2, 3):
    for dt in ("f64", "f32"):
        for nq in sorted({k[2] for k in best if k[0] == dim and k[1] == dt}):
            cells, entry = [], {}
            for be in ("rows", "pipe"):
                r = best.get((dim, dt, nq, be))
                if r:
                    cells.append(f"E={r['E']:>3s} T={r['threads']} R={r['R']} V={r.get('V') or 0} occ={r['ctas_per_sm']:>2s} {float(r['hbm_frac']):.3f}")
                    entry[be] = [int(r["E"]), int(r["threads"]), int(r["R"]), int(r.get("V") or 0)]
                    entry[be + "_frac"] = float(r["hbm_frac"])
                else:
                    cells.append("-")
            if "pipe" in entry:
                entry["prefer"] = "Pipe" if entry["pipe_frac"] >= entry.get("rows_frac", 0) else "Rows"
            over[f"{dim}:{dt}:{nq}"] = entry
            print(f"{dim}:{dt}:{nq:<5d} {cells[0]:34s} {cells[1]:34s} -> {entry.get('prefer','Rows')}")
if "--write" in sys.argv:
    path = os.path.join(HERE, "..", "rows_overrides.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(over)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)

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
_ovr = os.path.join(HERE, "..", "rows_overrides.json")
old_all = json.load(open(_ovr)) if os.path.exists(_ovr) else {}
print(f"{'case':12s} {'rows best':34s} {'pipe best':34s} {'mma best':34s}")
for dim in (2, 3):
    for dt in ("f64", "f32"):
        for nq in sorted({k[2] for k in best if k[0] == dim and k[1] == dt}):
            cells, entry = [], {}
            for be in ("rows", "pipe", "mma"):
                r = best.get((dim, dt, nq, be))
                if r:
                    cells.append(f"E={r['E']:>3s} T={r['threads']} R={r['R']} V={r.get('V') or 0} occ={r['ctas_per_sm']:>2s} {float(r['hbm_frac']):.3f}")
                    entry[be] = [int(r["E"]), int(r["threads"]), int(r["R"]), int(r.get("V") or 0)]
                    if be == "mma":  # (G, warps per CTA, MB0, NB1)
                        entry[be][1] //= 32
                    entry[be + "_frac"] = float(r["hbm_frac"])
                else:
                    cells.append("-")
            old_e = old_all.get(f"{dim}:{dt}:{nq}", {})
            merged = dict(old_e)
            for k, v in entry.items():
                if k.endswith("_frac") or k == "prefer":
                    continue
                if entry[k + "_frac"] >= merged.get(k + "_frac", 0):
                    merged[k], merged[k + "_frac"] = v, entry[k + "_frac"]
            fr = {b: merged.get(b + "_frac", 0) for b in ("rows", "pipe", "mma") if b in merged}
            if not merged.get("prefer_locked"):  # a hand-set preference (see "why") survives re-tuning
                merged["prefer"] = max(fr, key=fr.get).capitalize() if fr else "Rows"
            over[f"{dim}:{dt}:{nq}"] = merged
            print(f"{dim}:{dt}:{nq:<5d} {cells[0]:34s} {cells[1]:34s} {cells[2]:34s} -> {merged['prefer']}")
if "--write" in sys.argv:
    path = os.path.join(HERE, "..", "rows_overrides.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(over)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)

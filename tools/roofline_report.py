#!/usr/bin/env python
"""Roofline-aware sibling of the reference's postprocess.py; the original stays untouched.

Reads benchmark0N logs in the reference's three-lines-per-size format (`nelmt N Case: / norm: / DOF/s:` for
benchmark04/05, `Size N Case: / norm: / GB/s:` for benchmark01-03) together with the `info` side lines the drivers add
(`... HBM% of P GB/s, columns 6-11: ...`, `gpus G aggregate ...`), and prints one markdown table per log: per size the
best CUDA column, its fraction of the HBM roofline, the CPU column next to it and the ratio.  Several logs of the same
benchmark taken at different GPU counts (B200FE_NGPUS) are folded into a scaling table.  `--png` also draws the curves
when matplotlib is importable (it is not in the build image).

    python tools/roofline_report.py profiles/driver_logs/*.txt
"""
import argparse
import re
import sys


def parse(path):
    title, rows, info, gpus = None, {}, {}, 1
    for line in open(path, errors="replace"):
        t = line.split()
        if "NQ =" in line or (title is None and line.startswith("Benchmark")):
            title = line.strip()
        if len(t) > 3 and t[0] in ("nelmt", "Size") and t[2] in ("DOF/s:", "GB/s:"):
            rows[int(t[1])] = (t[2].rstrip(":"), [float(x) for x in t[3:]])
        elif t and t[0] == "info" and "HBM%" in line:
            m = re.search(r"gpus (\d+)", line)
            if m:
                gpus = int(m.group(1))
            tail = line.split(":", 1)[1].split("|")[0]
            vals = [float(x) for x in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", tail.split("(")[0])]
            info[int(t[1])] = (vals, "L2-resident" in line)
    return {"path": path, "title": title or path, "rows": rows, "info": info, "gpus": gpus}


def report(log, out=sys.stdout):
    out.write(f"\n### {log['title']}  ({log['path']}, {log['gpus']} GPU{'s' if log['gpus'] > 1 else ''})\n\n")
    out.write("| size | unit | best CUDA column | value | % of HBM roofline | CPU column 1 | GPU / CPU |\n|---|---|---|---|---|---|---|\n")
    for size in sorted(log["rows"]):
        unit, vals = log["rows"][size]
        first_cuda = 5 if unit == "DOF/s" else 2            # benchmark04/05: columns 6-11; benchmark01-03: columns 3-5
        cuda = vals[first_cuda:]
        if not cuda:
            continue
        k = max(range(len(cuda)), key=lambda i: cuda[i])
        pct, flag = "", ""
        if size in log["info"]:
            pv, small = log["info"][size]
            if len(pv) >= len(cuda):
                pct = f"{pv[len(pv) - len(cuda) + k]:.1f}"
            flag = " (L2-resident)" if small else ""
        cpu = vals[0]
        ratio = f"{cuda[k] / cpu:.0f}x" if cpu > 0 else "-"
        out.write(f"| {size} | {'G' if unit == 'DOF/s' else ''}{unit} | {first_cuda + k + 1} | {cuda[k]:.4g} | {pct}{flag} | {cpu:.4g} | {ratio} |\n")


def scaling(logs, out=sys.stdout):
    by = {}
    for lg in logs:
        by.setdefault(lg["title"], []).append(lg)
    for title, group in by.items():
        if len({g["gpus"] for g in group}) < 2:
            continue
        out.write(f"\n### scaling: {title}\n\n| GPUs | size | best value | per GPU | efficiency vs fewest GPUs |\n|---|---|---|---|---|\n")
        base = None
        for lg in sorted(group, key=lambda g: g["gpus"]):
            size = max(lg["rows"])
            unit, vals = lg["rows"][size]
            best = max(vals[5 if unit == "DOF/s" else 2:])
            per = best / lg["gpus"]
            base = base or per
            out.write(f"| {lg['gpus']} | {size} | {best:.4g} | {per:.4g} | {100 * per / base:.1f} % |\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("logs", nargs="+")
    ap.add_argument("--png", help="also draw GDoF/s (or GB/s) against size into this file (needs matplotlib)")
    args = ap.parse_args()
    logs = [parse(p) for p in args.logs]
    logs = [lg for lg in logs if lg["rows"]]
    for lg in logs:
        report(lg)
    scaling(logs)
    if args.png:
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except ImportError:
            sys.exit("matplotlib is not installed: tables only")
        for lg in logs:
            xs = sorted(lg["rows"])
            unit = lg["rows"][xs[0]][0]
            first = 5 if unit == "DOF/s" else 2
            plt.loglog(xs, [max(lg["rows"][x][1][first:]) for x in xs], label=f"{lg['title']} ({lg['gpus']} GPU)")
        plt.xlabel("size")
        plt.ylabel("best CUDA column")
        plt.legend(fontsize=6)
        plt.savefig(args.png, dpi=150)
    return 0


if __name__ == "__main__":
    sys.exit(main())

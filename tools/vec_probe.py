#!/usr/bin/env python3
"""benchmark01-03 kernels at a few large sizes, ours against the reference's kernels (oracle/_ref): GB/s, device-timed"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200fe_loader
from tools import bench_sweeps as bs

fe = b200fe_loader.load()
v = bs.sweep_vec(fe, torch, 6546.9, same_box=bs.SameBox(), max_log2=int(os.environ.get("MAXLOG2", 30)), reps=8)
for b in ("b01", "b02", "b03"):
    for r in v[b]:
        if r["log2"] >= 24:
            print(b, r.get("size", r["n"]), "vl", r["vl"]["gb_s"], "scalar", r["scalar"]["gb_s"], "ref", r.get("ref_kernel_gb_s"))

#!/usr/bin/env python3
"""Times the headline operator (hex BwdTrans nq = 8 FP64, 262 144 elements) in bursts of 40 launches: mean / min launch
time and fraction of the measured HBM roofline.  Experiment knobs are read by the library from the environment (e.g.
B200FE_MMA_SCHED), so run one process per setting:  B200FE_MMA_SCHED=1 python tools/headline_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200fe_loader

fe = b200fe_loader.load()
nq, nm, nelmt = 8, 7, int(os.environ.get("NELMT", 262144))
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6546.9
st = torch.cuda.current_stream().cuda_stream
b = torch.cos(torch.arange(nm * nq, dtype=torch.float64)).cuda()
d_in = torch.randn(nelmt * nm ** 3, dtype=torch.float64, device="cuda")
d_out = torch.empty(nelmt * nq ** 3, dtype=torch.float64, device="cuda")
if len(sys.argv) > 1:
    fe.set_backend(sys.argv[1])


def step():
    fe.bwdtrans_hex("BwdTransHexKernel_QP_Shared", "f64", nq, nq, nq, nelmt, b.data_ptr(), b.data_ptr(), b.data_ptr(),
                    d_in.data_ptr(), d_out.data_ptr(), stream=st)


for _ in range(5):
    step()
torch.cuda.synchronize()
byts = nelmt * 8 * (nm ** 3 + nq ** 3)
res = []
for burst in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    ev[0].record()
    for k in range(40):
        step()
        ev[k + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(40))
    res.append((sum(ts) / 40, ts[0]))
    import time
    time.sleep(0.5)
mean = min(r[0] for r in res)
best = min(r[1] for r in res)
print(f"sched={os.environ.get('B200FE_MMA_SCHED', '0')} backend={fe.last_backend()} mean {mean:.4f} ms "
      f"({1e-6 * byts / mean / peak:.4f}) min {best:.4f} ms ({1e-6 * byts / best / peak:.4f})")

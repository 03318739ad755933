#!/usr/bin/env python3
"""Does a bandwidth-bound kernel slow down under sustained load on this B200, and which clock moves?
Runs (a) a plain device copy and (b) the hex nq=8 FP64 operator back to back for ~150 ms each, records the
per-launch time series (CUDA events) and NVML SM/memory clocks + power every millisecond."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pynvml

import b200fe_loader

fe = b200fe_loader.load()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], threading.Event()


def sampler():
    t0 = time.perf_counter()
    while not stop.is_set():
        try:
            samples.append((time.perf_counter() - t0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                            pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        except Exception as e:  # noqa
            samples.append((time.perf_counter() - t0, -1, -1, -1, str(e)))
        time.sleep(0.001)


def series(fn, n):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for k in range(n):
        fn()
        ev[k + 1].record()
    torch.cuda.synchronize()
    return [ev[k].elapsed_time(ev[k + 1]) for k in range(n)]


def show(name, ts, bytes_per):
    ts = np.array(ts)
    print(f"{name}: first10 {ts[:10].mean():.4f} ms, last10 {ts[-10:].mean():.4f} ms, min {ts.min():.4f}, "
          f"GB/s first {bytes_per / ts[:10].mean() / 1e6:.0f} last {bytes_per / ts[-10:].mean() / 1e6:.0f}")
    print("   every 10th:", " ".join(f"{t:.3f}" for t in ts[::10]))


nq, nm, nelmt = 8, 7, 262144
a = torch.empty(1 << 27, dtype=torch.float64, device="cuda").normal_()
b = torch.empty_like(a)
d_b = torch.cos(torch.arange(nm * nq, dtype=torch.float64)).cuda()
d_in = torch.randn(nelmt * nm ** 3, dtype=torch.float64, device="cuda")
d_out = torch.empty(nelmt * nq ** 3, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def op():
    fe.bwdtrans_hex("BwdTransHexKernel_QP_Shared", "f64", nq, nq, nq, nelmt, d_b.data_ptr(), d_b.data_ptr(),
                    d_b.data_ptr(), d_in.data_ptr(), d_out.data_ptr(), stream=st)


thr = threading.Thread(target=sampler, daemon=True)
thr.start()
time.sleep(0.05)
for name, fn, n, nbytes in (("copy 1 GiB+1 GiB", lambda: b.copy_(a), 400, 2 * a.numel() * 8),
                            ("hex8 f64 262144", op, 500, nelmt * 8 * (nm ** 3 + nq ** 3))):
    for _ in range(3):
        fn()
    time.sleep(0.5)  # idle: let the GPU cool / clocks settle
    t_mark = time.perf_counter()
    ts = series(fn, n)
    show(name, ts, nbytes)
    time.sleep(0.3)
stop.set()
thr.join()
print("t_s sm_mhz mem_mhz power_w reasons  (every 8th sample)")
for s in samples[::8]:
    print(f"{s[0]:.3f} {s[1]} {s[2]} {s[3]:.0f} {s[4]}")

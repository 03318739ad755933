#!/usr/bin/env python3
"""quad FP32 nq = 32 through the tcgen05 back-end ("umma") against the oracle in double: component-wise error, timing"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200fe_loader
import oracle

fe = b200fe_loader.load()
st = torch.cuda.current_stream().cuda_stream
nq, nm = 32, 31
rng = np.random.default_rng(1)
backend = sys.argv[1] if len(sys.argv) > 1 else "umma"
for nelmt in (4, 8, 64, 4096, 4097, 3, 1, 601):
    b0 = rng.standard_normal(nm * nq).astype(np.float32)
    b1 = rng.standard_normal(nm * nq).astype(np.float32)
    inp = rng.standard_normal(nelmt * nm * nm).astype(np.float32)
    d_b0, d_b1, d_in = (torch.from_numpy(a).cuda() for a in (b0, b1, inp))
    d_out = torch.full((nelmt * nq * nq,), float("nan"), dtype=torch.float32, device="cuda")
    fe.set_backend(backend)
    fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f32", nq, nq, nelmt, d_b0.data_ptr(), d_b1.data_ptr(), d_in.data_ptr(),
                     d_out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().astype(np.float64)
    want = oracle.bwdtrans_quad(nq, nq, nelmt, b0.astype(np.float64), b1.astype(np.float64), inp.astype(np.float64))
    scale = oracle.bwdtrans_quad(nq, nq, nelmt, np.abs(b0).astype(np.float64), np.abs(b1).astype(np.float64),
                                 np.abs(inp).astype(np.float64))
    err = np.abs(got - want)
    bad = np.argwhere(~(err <= 1e-5 * scale)).ravel()
    print(f"nelmt={nelmt:5d} backend={fe.last_backend()} max componentwise err {np.nanmax(err / scale):.3e} "
          f"normwise {np.nanmax(err) / np.abs(want).max():.3e} nan={int(np.isnan(got).sum())} bad={bad.size}",
          (bad[:6], got[bad[:3]], want[bad[:3]]) if bad.size else "")
# timing at the BASELINE size
nelmt = 65536
d_b = torch.from_numpy(np.cos(np.arange(nm * nq, dtype=np.float32))).cuda()
d_in = torch.randn(nelmt * nm * nm, dtype=torch.float32, device="cuda")
d_out = torch.empty(nelmt * nq * nq, dtype=torch.float32, device="cuda")
for be in (backend, "mma"):
    fe.set_backend(be)
    call = lambda: fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f32", nq, nq, nelmt, d_b.data_ptr(), d_b.data_ptr(),
                                    d_in.data_ptr(), d_out.data_ptr(), stream=st)
    for _ in range(3):
        call()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for k in range(10):
        call()
        ev[k + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[k].elapsed_time(ev[k + 1]) for k in range(10))
    byts = nelmt * 4 * (nm * nm + nq * nq)
    print(f"{be}: {ms:.4f} ms  {1e-6 * byts / ms:.0f} GB/s  frac {1e-6 * byts / ms / 6546.9:.3f}  "
          f"{1e-6 * nelmt * nm * nm / ms:.1f} GDoF/s")
if os.environ.get("PROF"):
    os.environ["B200FE_UMMA_PROF"] = "1"
    fe.set_backend("umma")
    fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f32", nq, nq, nelmt, d_b.data_ptr(), d_b.data_ptr(),
                     d_in.data_ptr(), d_out.data_ptr(), stream=st)
    torch.cuda.synchronize()

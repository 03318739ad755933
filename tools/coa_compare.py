#!/usr/bin/env python
"""Time every back-end that serves the interleaved (`_Coa`) entry points, per (operator, nq, dtype), at ~64 Mi
quadrature points through the C ABI -- the measurement behind the default routing in csrc/bwdtrans_{quad,hex}.cu.

    python tools/coa_compare.py > gpurun_out/coa_compare.csv
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200fe_loader

fe = b200fe_loader.load()
PEAK = 6546.9


def main():
    st = torch.cuda.current_stream().cuda_stream
    print("op,nq,dtype,backend,ms,gb_s,hbm_frac")
    for dim, nqs, kern in ((2, (3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 24, 32), "BwdTransQuadKernel_Coa"),
                           (3, (3, 4, 5, 6, 7, 8, 9, 10, 12), "BwdTransHexKernel_Coa")):
        for suf, tdt, size in (("f64", torch.float64, 8), ("f32", torch.float32, 4)):
            for nq in nqs:
                nm = nq - 1
                nelmt = max(32, ((1 << 26) // nq ** dim) // 32 * 32)
                rng = np.random.default_rng(nq)
                b = torch.from_numpy(rng.standard_normal(nm * nq)).to(tdt).cuda()
                d_in = torch.randn(nelmt * nm ** dim, dtype=tdt, device="cuda")
                d_out = torch.empty(nelmt * nq ** dim, dtype=tdt, device="cuda")
                first = None
                for be in ("tpe", "lanes", "rows", "auto"):
                    fe.set_backend(be)

                    def call():
                        if dim == 2:
                            fe.bwdtrans_quad(kern, suf, nq, nq, nelmt, b.data_ptr(), b.data_ptr(), d_in.data_ptr(),
                                             d_out.data_ptr(), stream=st)
                        else:
                            fe.bwdtrans_hex(kern, suf, nq, nq, nq, nelmt, b.data_ptr(), b.data_ptr(), b.data_ptr(),
                                            d_in.data_ptr(), d_out.data_ptr(), stream=st)
                    try:
                        d_out.fill_(float("nan"))
                        call()
                    except Exception:
                        continue
                    torch.cuda.synchronize()
                    if first is None:
                        first = d_out.clone()
                    same = bool(torch.equal(first, d_out))
                    reps = 6
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
                    for r in range(reps):
                        ev[2 * r].record()
                        call()
                        ev[2 * r + 1].record()
                    torch.cuda.synchronize()
                    ms = min(ev[2 * r].elapsed_time(ev[2 * r + 1]) for r in range(1, reps))
                    gbs = 1e-9 * nelmt * (nm ** dim + nq ** dim) * size / (ms * 1e-3)
                    print(f"{'quad' if dim == 2 else 'hex'},{nq},{suf},{be}:{fe.last_backend()},{ms:.4f},{gbs:.1f},"
                          f"{gbs / PEAK:.3f}{'' if same else ',MISMATCH'}", flush=True)
                fe.set_backend("auto")
                del d_in, d_out


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-call entry points against b200fe_plan_* (basis staged once) at ~64 Mi quadrature points: mean time of a call
in a train of back-to-back calls on one stream, the way run_test's 40 repetitions or a solver's time steps issue them.

    python tools/plan_probe.py > gpurun_out/plan_probe.csv
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200fe_loader

fe = b200fe_loader.load()
PEAK = 6546.9
TRAIN = 20


def train_ms(call):
    best = None
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(TRAIN):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / TRAIN
        best = ms if best is None else min(best, ms)
    return best


def main():
    st = torch.cuda.current_stream().cuda_stream
    print("op,layout,nq,dtype,backend,per_call_ms,plan_ms,plan_gain,per_call_hbm_frac,plan_hbm_frac,same_bits")
    for dim, nqs in ((2, (2, 4, 6, 8, 10, 12, 14, 16, 32)), (3, (2, 4, 6, 8, 10))):
        for suf, tdt, size in (("f64", torch.float64, 8), ("f32", torch.float32, 4)):
            for coa in (False, True):
                for nq in nqs:
                    nm = nq - 1
                    nelmt = max(32, ((1 << 26) // nq ** dim) // 32 * 32)
                    rng = np.random.default_rng(nq)
                    b = torch.from_numpy(rng.standard_normal(nm * nq)).to(tdt).cuda()
                    d_in = torch.randn(nelmt * nm ** dim, dtype=tdt, device="cuda")
                    d_out = torch.empty(nelmt * nq ** dim, dtype=tdt, device="cuda")
                    kern = ("BwdTransQuadKernel" if dim == 2 else "BwdTransHexKernel") + ("_Coa" if coa else "_QP_Shared")

                    def per_call():
                        if dim == 2:
                            fe.bwdtrans_quad(kern, suf, nq, nq, nelmt, b.data_ptr(), b.data_ptr(), d_in.data_ptr(),
                                             d_out.data_ptr(), stream=st)
                        else:
                            fe.bwdtrans_hex(kern, suf, nq, nq, nq, nelmt, b.data_ptr(), b.data_ptr(), b.data_ptr(),
                                            d_in.data_ptr(), d_out.data_ptr(), stream=st)

                    plan = fe.Plan(dim, suf, nq, [b.data_ptr()] * dim, stream=st)

                    def planned():
                        plan.bwdtrans(nelmt, d_in.data_ptr(), d_out.data_ptr(), coa=coa, stream=st)

                    per_call()
                    torch.cuda.synchronize()
                    first = d_out.clone()
                    d_out.fill_(float("nan"))
                    planned()
                    torch.cuda.synchronize()
                    same = bool(torch.equal(first, d_out))
                    del first
                    t_call = train_ms(per_call)
                    t_plan = train_ms(planned)
                    byts = nelmt * (nm ** dim + nq ** dim) * size
                    f_call, f_plan = (1e-9 * byts / (t * 1e-3) / PEAK for t in (t_call, t_plan))
                    print(f"{'quad' if dim == 2 else 'hex'},{'coa' if coa else 'em'},{nq},{suf},{fe.last_backend()},"
                          f"{t_call:.4f},{t_plan:.4f},{t_call / t_plan:.3f},{f_call:.3f},{f_plan:.3f},{same}", flush=True)
                    plan.destroy()
                    del d_in, d_out


if __name__ == "__main__":
    main()

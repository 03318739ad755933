#!/usr/bin/env python
"""IProductWRTBase: the lanes kernel (sumfac_iprod_lanes.cuh) against the row / tensor-core kernels, per (operator, nq, dtype), weighted and unweighted, at ~64 Mi quadrature points through the C ABI; every
variant must store exactly the bits of the row kernel.

    python tools/ipl_probe.py > gpurun_out/ipl_probe.csv
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200fe_loader

fe = b200fe_loader.load()
PEAK = 6546.9


def main():
    st = torch.cuda.current_stream().cuda_stream
    print("op,nq,dtype,weighted,variant,backend,ms,hbm_frac,same_bits")
    only_hex = len(sys.argv) > 1 and sys.argv[1] == "hex"
    for dim, nqs in ((3, (8, 10)),) if only_hex else ((2, (4, 6, 8, 10, 12, 14, 16)), (3, (4, 6, 8, 10))):
        for suf, tdt, size in (("f64", torch.float64, 8), ("f32", torch.float32, 4)):
            for nq in nqs:
                nm = nq - 1
                nelmt = ((1 << 26) // nq ** dim) // 32 * 32 - 5
                b = torch.randn(nm * nq, dtype=tdt, device="cuda")
                x = torch.randn(nelmt * nq ** dim, dtype=tdt, device="cuda")
                wgt = torch.rand(nelmt * nq ** dim, dtype=tdt, device="cuda") + 0.5
                y = torch.empty(nelmt * nm ** dim, dtype=tdt, device="cuda")
                for weighted in (0, 1):
                    ref = None
                    variants = [("rows", None), ("mma", None), ("lanes", None), ("pipe", None)]  # pipe: hexes nq = 8, 10  # tile sizes: tools/tune/iprod_probe.cu
                    for be, el in variants:
                        fe.set_backend(be)
                        def call():
                            fe.iproduct(suf, (nq,) * dim, nelmt, [b.data_ptr()] * dim, x.data_ptr(), y.data_ptr(),
                                        weights=wgt.data_ptr() if weighted else 0, stream=st)
                        try:
                            y.fill_(float("nan"))
                            call()
                            torch.cuda.synchronize()
                        except Exception:
                            continue
                        if ref is None:
                            ref = y.clone()
                        same = bool(torch.equal(ref, y))
                        ts = []
                        for _ in range(6):
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record()
                            call()
                            e1.record()
                            torch.cuda.synchronize()
                            ts.append(e0.elapsed_time(e1))
                        ms = min(ts[1:])
                        byts = nelmt * size * ((1 + weighted) * nq ** dim + nm ** dim)
                        print(f"{'quad' if dim == 2 else 'hex'},{nq},{suf},{weighted},{be}{el or ''},{fe.last_backend()},"
                              f"{ms:.4f},{1e-9 * byts / (ms * 1e-3) / PEAK:.3f},{same}", flush=True)
                    fe.set_backend("auto")
                del x, wgt, y


if __name__ == "__main__":
    main()

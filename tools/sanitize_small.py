#!/usr/bin/env python3
"""Small-size runs of every back-end for compute-sanitizer (memcheck): python tools/sanitize_small.py
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import b200fe_loader

fe = b200fe_loader.load()
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(0)
n = 0
for suf, dt in (("f64", torch.float64), ("f32", torch.float32)):
    for be in ("rows", "pipe", "mma", "nm1", "generic", "auto"):
        for dim, nqs in ((2, (2, 4, 8, 12, 14, 16, 32)), (3, (2, 4, 6, 8, 10))):
            for nq in nqs:
                nm = nq - 1
                for nelmt in (1, 37, 300):
                    b = torch.randn(nm * nq, dtype=dt, device="cuda")
                    x = torch.randn(nelmt * nm ** dim, dtype=dt, device="cuda")
                    y = torch.empty(nelmt * nq ** dim, dtype=dt, device="cuda")
                    try:
                        fe.set_backend(be)
                        if dim == 2:
                            fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", suf, nq, nq, nelmt, b.data_ptr(), b.data_ptr(),
                                             x.data_ptr(), y.data_ptr(), stream=st)
                        else:
                            fe.bwdtrans_hex("BwdTransHexKernel_QP_Shared", suf, nq, nq, nq, nelmt, b.data_ptr(),
                                            b.data_ptr(), b.data_ptr(), x.data_ptr(), y.data_ptr(), stream=st)
                        n += 1
                    except fe.B200feError as e:
                        assert e.code == fe.E_UNSUPPORTED, e
                    finally:
                        fe.set_backend("auto")
                    torch.cuda.synchronize()
                    assert torch.isfinite(y).all()
print("launched", n, "operator calls, all outputs finite")

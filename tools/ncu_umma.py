#!/usr/bin/env python3
"""A few launches of quad BwdTrans FP32 nq = 32 at the BASELINE size (65 536 elements) through the default route (the
tcgen05 back-end) for `ncu --set full -k regex:umma` (tools/gpu_r2_b.sh)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200fe_loader

fe = b200fe_loader.load()
st = torch.cuda.current_stream().cuda_stream
nq, nm, nelmt = 32, 31, 65536
b = torch.cos(torch.arange(nm * nq, dtype=torch.float32)).cuda()
d_in = torch.randn(nelmt * nm * nm, dtype=torch.float32, device="cuda")
d_out = torch.empty(nelmt * nq * nq, dtype=torch.float32, device="cuda")
for _ in range(3):
    fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f32", nq, nq, nelmt, b.data_ptr(), b.data_ptr(), d_in.data_ptr(),
                     d_out.data_ptr(), stream=st)
torch.cuda.synchronize()
print("ok", fe.last_backend(), float(d_out[:16].abs().sum().item()))

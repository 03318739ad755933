#!/usr/bin/env python3
"""One launch each of the benchmark01-03 kernels at 2^28 doubles (2 GiB: far beyond L2), for `ncu --set full`
(tools/gpu_r2_a.sh): reduce_partials_kernel (l2norm_vl + reduce_vl), add_vector_kernel, matvec_kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import b200fe_loader

fe = b200fe_loader.load()
st = torch.cuda.current_stream().cuda_stream
n = 1 << 28
x = torch.empty(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")
fe.set_data("f64", x.data_ptr(), n, stream=st)
fe.set_data("f64", y.data_ptr(), n, second=True, stream=st)
sums = torch.zeros(1024, dtype=torch.float64, device="cuda")
res = torch.zeros(1, dtype=torch.float64, device="cuda")
fe.l2norm_vl("f64", sums.data_ptr(), x.data_ptr(), n, 1024, 1, stream=st)
fe.reduce_vl("f64", res.data_ptr(), sums.data_ptr(), 1024, 1, stream=st)
fe.add_vector("f64", x.data_ptr(), y.data_ptr(), n, 1, stream=st)
size = 16384
xv = torch.arange(size, dtype=torch.float64, device="cuda")
yv = torch.empty(size, dtype=torch.float64, device="cuda")
fe.compute_matvec("f64", size, size, y.data_ptr(), xv.data_ptr(), yv.data_ptr(), 1, stream=st)
torch.cuda.synchronize()
print("ok", float(res.item()) ** 0.5, float(yv[0].item()))

"""helpers shared by the -m gpu tests: device buffers through torch, calls through the C ABI"""
import numpy as np
import torch

import b200fe_loader

fe = b200fe_loader.load()

NP = {"f64": np.float64, "f32": np.float32}
TOL = {"f64": 1e-12, "f32": 1e-5}  # north_star tolerances (relative)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def rel_max(got, want):
    scale = float(np.abs(want).max()) or 1.0
    return float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) / scale


def run_quad(kernel, suf, nq0, nq1, nelmt, b0, b1, inp, nm0=None, nm1=None):
    nm0 = nq0 - 1 if nm0 is None else nm0
    nm1 = nq1 - 1 if nm1 is None else nm1
    d_b0, d_b1, d_in = dev(b0), dev(b1), dev(inp)
    d_out = torch.full((nelmt * nq0 * nq1,), float("nan"), dtype=d_in.dtype, device="cuda")
    d_wsp = torch.empty(max(1, nelmt * nq0 * nm1), dtype=d_in.dtype, device="cuda")
    fe.bwdtrans_quad(kernel, suf, nq0, nq1, nelmt, d_b0.data_ptr(), d_b1.data_ptr(), d_in.data_ptr(),
                     d_out.data_ptr(), wsp=d_wsp.data_ptr(), stream=torch.cuda.current_stream().cuda_stream,
                     nm0=nm0, nm1=nm1)
    return host(d_out)


def run_hex(kernel, suf, nq, nelmt, b, inp, nm=None):
    nq0, nq1, nq2 = nq
    nm = (nq0 - 1, nq1 - 1, nq2 - 1) if nm is None else nm
    d_b = [dev(x) for x in b]
    d_in = dev(inp)
    d_out = torch.full((nelmt * nq0 * nq1 * nq2,), float("nan"), dtype=d_in.dtype, device="cuda")
    d_w0 = torch.empty(max(1, nelmt * nq0 * nm[1] * nm[2]), dtype=d_in.dtype, device="cuda")
    d_w1 = torch.empty(max(1, nelmt * nq0 * nq1 * nm[2]), dtype=d_in.dtype, device="cuda")
    fe.bwdtrans_hex(kernel, suf, nq0, nq1, nq2, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(), d_b[2].data_ptr(),
                    d_in.data_ptr(), d_out.data_ptr(), wsp0=d_w0.data_ptr(), wsp1=d_w1.data_ptr(),
                    stream=torch.cuda.current_stream().cuda_stream, nm=nm)
    return host(d_out)


def assert_parity(got, want, suf, what=""):
    """Bit for bit with the oracle / the reference kernels -- every back-end accumulates in the reference's own
    order with fused multiply-adds -- except the FP32 tensor-core back-ends (3xTF32 split: "mma" = warp-level
    mma.sync, "umma" = tcgen05, "coa-mma" = mma.sync with M = elements in the interleaved layout), which agree to rounding and are held to north_star's FP32 tolerance here (norm-wise)
    and to the component-wise bound in tests/test_fullsize_gpu.py / tests/test_umma_gpu.py."""
    if suf == "f32" and fe.last_backend() in ("mma", "umma", "coa-mma"):
        err = rel_max(got, want)
        assert err < TOL["f32"], (what, fe.last_backend() + " f32", err)
    else:
        assert np.array_equal(got, want), (what, fe.last_backend(), rel_max(got, want))

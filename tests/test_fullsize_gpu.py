"""Full BASELINE-size parity on ELEMENT-DEPENDENT data (every element carries different modes), bit for bit:

  * hex nq = 8 FP64 at 131 072 elements (configs[3]: 64 Mi quadrature points) and 262 144 elements (bench.py's per-GPU
    shard), quad nq = 4 FP64 at 4 194 304 elements (configs[2]) -- libb200fe through the C ABI against the CPU oracle
    (oracle/oracle_impl.h, the restatement of benchmark05.cc:57-101 / benchmark04.cc:49-72) AND against the
    reference's own kernel compiled for sm_100a (oracle/_ref, `QP/Shared` variant: benchmark05.cc:291-429,
    benchmark04.cc:206-300), both layouts;
  * the FP32 tensor-core route (3xTF32, the default for quad nq = 32): a COMPONENT-WISE bound.  Each output is a double
    sum of products, so its rounding error is bounded relative to the same sum taken over absolute values,
    |got - want| <= tol * (|B1|^T (|B0|^T |in|)) -- the classical componentwise bound for dot products.  That is what the
    reference's own FFMA chain satisfies with tol ~ nm * 2^-24; the tensor-core route is held to north_star's 1e-5
    against it, element by element, with no division by max|want| (an output near a cancellation cannot hide).
"""
import ctypes
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so")


@pytest.fixture(scope="module")
def G():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests import gpu_util
    return gpu_util


@pytest.fixture(scope="module")
def ref():
    return ctypes.CDLL(REF_SO) if os.path.exists(REF_SO) else None


def vp(t):
    return ctypes.c_void_p(t.data_ptr())


def u(x):
    return ctypes.c_uint(int(x))


def element_dependent(torch, n, dtype, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.randn(n, dtype=dtype, device="cuda", generator=g)


@pytest.mark.parametrize("nelmt", [131072, 262144])
def test_hex_nq8_full_size_element_dependent_bit_exact(G, ref, nelmt):
    import torch
    nq, nm = 8, 7
    rng = np.random.default_rng(8100 + nelmt % 97)
    b = [rng.standard_normal(nm * nq) for _ in range(3)]
    d_b = [G.dev(x) for x in b]
    d_in = element_dependent(torch, nelmt * nm ** 3, torch.float64, 5 + nelmt)
    d_out = torch.full((nelmt * nq ** 3,), float("nan"), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    G.fe.bwdtrans_hex("BwdTransHexKernel_QP_Shared", "f64", nq, nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                      d_b[2].data_ptr(), d_in.data_ptr(), d_out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert G.fe.last_backend() == "mma"          # the headline kernel
    # (1) the CPU oracle, every one of the nelmt * 512 outputs
    want = oracle.bwdtrans_hex(nq, nq, nq, nelmt, b[0], b[1], b[2], d_in.cpu().numpy(), use_fma=True)
    got = d_out.cpu().numpy()
    assert np.array_equal(got, want), float(np.abs(got - want).max())
    del got
    # (2) the reference's own kernel on the same device
    if ref is not None:
        r_out = torch.full_like(d_out, float("nan"))
        rc = ref.ref_bwdtrans_hex_f64(ctypes.c_int(3), u(nq), u(nq), u(nq), u(nelmt), vp(d_b[0]), vp(d_b[1]), vp(d_b[2]),
                                      vp(d_in), None, None, None, None, vp(r_out), u(128), u(1), ctypes.c_void_p(st))
        assert rc == 0
        torch.cuda.synchronize()
        assert torch.equal(d_out, r_out)
        del r_out
    # (3) the interleaved entry point on the same data, re-laid out
    want_coa = oracle.to_coa(want, nelmt, nq ** 3)
    d_in_coa = G.dev(oracle.to_coa(d_in.cpu().numpy(), nelmt, nm ** 3))
    d_out.fill_(float("nan"))
    G.fe.bwdtrans_hex("BwdTransHexKernel_Coa", "f64", nq, nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                      d_b[2].data_ptr(), d_in_coa.data_ptr(), d_out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), want_coa)


def test_quad_nq4_full_size_element_dependent_bit_exact(G, ref):
    import torch
    nq, nm, nelmt = 4, 3, 4194304
    rng = np.random.default_rng(4400)
    b = [rng.standard_normal(nm * nq) for _ in range(2)]
    d_b = [G.dev(x) for x in b]
    d_in = element_dependent(torch, nelmt * nm * nm, torch.float64, 44)
    d_out = torch.full((nelmt * nq * nq,), float("nan"), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f64", nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                       d_in.data_ptr(), d_out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    h_in = d_in.cpu().numpy()
    want = oracle.bwdtrans_quad(nq, nq, nelmt, b[0], b[1], h_in, use_fma=True)
    assert np.array_equal(d_out.cpu().numpy(), want)
    if ref is not None:
        r_out = torch.full_like(d_out, float("nan"))
        rc = ref.ref_bwdtrans_quad_f64(ctypes.c_int(3), u(nq), u(nq), u(nelmt), vp(d_b[0]), vp(d_b[1]), vp(d_in), None,
                                       None, vp(r_out), u(128), u(1), ctypes.c_void_p(st))
        assert rc == 0
        torch.cuda.synchronize()
        assert torch.equal(d_out, r_out)
        del r_out
    d_in_coa = G.dev(oracle.to_coa(h_in, nelmt, nm * nm))
    d_out.fill_(float("nan"))
    G.fe.bwdtrans_quad("BwdTransQuadKernel_Coa", "f64", nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                       d_in_coa.data_ptr(), d_out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), oracle.to_coa(want, nelmt, nq * nq))


@pytest.mark.parametrize("nq,backend", [(16, "mma"), (32, "mma"), (32, "umma")])
@pytest.mark.parametrize("kind", ["random", "cancelling"])
def test_fp32_tensor_core_route_componentwise_bound(G, nq, backend, kind):
    """quad FP32 on the tensor cores (3xTF32 split; "mma": warp-level mma.sync, "umma": tcgen05 + TMEM, the default at
    nq = 32): |got - want| <= 1e-5 * (|B1|^T |B0|^T |in|) for EVERY output"""
    nm, nelmt = nq - 1, 4096
    rng = np.random.default_rng(3200 + nq)
    b0 = rng.standard_normal(nm * nq).astype(np.float32)
    b1 = rng.standard_normal(nm * nq).astype(np.float32)
    inp = rng.standard_normal(nelmt * nm * nm).astype(np.float32)
    if kind == "cancelling":
        # the reference's synthetic data: cos/sin tables whose sums cancel heavily (outputs far below the operand scale)
        b0 = oracle.gen_basis(nm, nq, np.float32)
        b1 = b0.copy()
        inp = oracle.gen_in(nelmt, nm * nm, np.float32) * (1.0 + 1e-3 * rng.standard_normal(nelmt * nm * nm)).astype(np.float32)
    G.fe.set_backend(backend)
    try:
        got = G.run_quad("BwdTransQuadKernel_QP_Shared", "f32", nq, nq, nelmt, b0, b1, inp)
        assert G.fe.last_backend() == backend
    finally:
        G.fe.set_backend("auto")
    # exact-ish reference and the componentwise scale, both in double
    want = oracle.bwdtrans_quad(nq, nq, nelmt, b0.astype(np.float64), b1.astype(np.float64), inp.astype(np.float64))
    scale = oracle.bwdtrans_quad(nq, nq, nelmt, np.abs(b0).astype(np.float64), np.abs(b1).astype(np.float64),
                                 np.abs(inp).astype(np.float64))
    err = np.abs(got.astype(np.float64) - want)
    worst = float((err / scale).max())
    assert worst < 1e-5, (nq, kind, worst)
    # and the reference's own FP32 arithmetic (FFMA chain, bit-exact back-end) sits under the same bound
    chain = oracle.bwdtrans_quad(nq, nq, nelmt, b0, b1, inp, use_fma=True)
    assert float((np.abs(chain.astype(np.float64) - want) / scale).max()) < 1e-5

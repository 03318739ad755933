"""BwdTrans in its GEMM formulation on the library's own kernels (b200fe_gemm_bwdtrans_*: the factorisation
the reference gives cuBLAS, benchmark04.cc:804-820 / benchmark05.cc:1128-1153, intermediates in global memory) and the
batched small mat-vec (b200fe_matvec_batched_*, benchmark03's operator benchmark03.cc:80-104 over many matrices).

The GEMM formulation sums in the reference kernels' order with fused multiply-adds, so it must be BIT-IDENTICAL to the
oracle and to the fused entry points; against the reference's cuBLAS formulation (oracle/libref_blas.so) it agrees to
rounding.  The batched mat-vec is compared with the oracle's long-double row dots (1e-12 / 1e-5 of the row's absolute
sum) and must be deterministic."""
import ctypes
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLAS_SO = os.path.join(ROOT, "oracle", "libref_blas.so")


@pytest.fixture(scope="module")
def G():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests import gpu_util
    return gpu_util


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("nq", [(2, 2), (4, 4), (7, 7), (8, 8), (16, 16), (32, 32), (5, 9)])
def test_quad_gemm_formulation_bit_identical_to_the_fused_operator(G, suf, nq):
    import torch
    dt, nelmt = G.NP[suf], 333
    nm = (nq[0] - 1, nq[1] - 1)
    rng = np.random.default_rng(40 + nq[0])
    b0, b1 = rng.standard_normal(nm[0] * nq[0]).astype(dt), rng.standard_normal(nm[1] * nq[1]).astype(dt)
    inp = rng.standard_normal(nelmt * nm[0] * nm[1]).astype(dt)
    d_b0, d_b1, d_in = G.dev(b0), G.dev(b1), G.dev(inp)
    d_out = torch.full((nelmt * nq[0] * nq[1],), float("nan"), dtype=d_in.dtype, device="cuda")
    d_wsp = torch.full((nelmt * nq[0] * nm[1],), float("nan"), dtype=d_in.dtype, device="cuda")
    G.fe.gemm_bwdtrans(suf, nq, nelmt, [d_b0.data_ptr(), d_b1.data_ptr()], d_in.data_ptr(), d_out.data_ptr(),
                       [d_wsp.data_ptr()], stream=torch.cuda.current_stream().cuda_stream)
    got = G.host(d_out)
    want = oracle.bwdtrans_quad(nq[0], nq[1], nelmt, b0, b1, inp, use_fma=True)
    assert np.array_equal(got, want), G.rel_max(got, want)
    # the intermediate is the reference's wsp[(e,q)][i] (benchmark04.cc:161-177)
    wsp = G.host(d_wsp).reshape(nelmt * nm[1], nq[0]).astype(np.float64)
    ref = inp.reshape(nelmt * nm[1], nm[0]).astype(np.float64) @ b0.reshape(nm[0], nq[0]).astype(np.float64)
    assert np.abs(wsp - ref).max() <= (1e-12 if suf == "f64" else 1e-5) * np.abs(ref).max()


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("nq", [(2, 2, 2), (4, 4, 4), (8, 8, 8), (10, 10, 10), (3, 4, 5)])
def test_hex_gemm_formulation_bit_identical_and_close_to_cublas(G, suf, nq):
    import torch
    dt, nelmt = G.NP[suf], 97
    nm = tuple(n - 1 for n in nq)
    rng = np.random.default_rng(50 + nq[0])
    b = [rng.standard_normal(nm[d] * nq[d]).astype(dt) for d in range(3)]
    inp = rng.standard_normal(nelmt * nm[0] * nm[1] * nm[2]).astype(dt)
    d_b, d_in = [G.dev(x) for x in b], G.dev(inp)
    d_out = torch.full((nelmt * nq[0] * nq[1] * nq[2],), float("nan"), dtype=d_in.dtype, device="cuda")
    w1 = torch.empty(nelmt * nq[0] * nm[1] * nm[2], dtype=d_in.dtype, device="cuda")
    w2 = torch.empty(nelmt * nq[0] * nq[1] * nm[2], dtype=d_in.dtype, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    G.fe.gemm_bwdtrans(suf, nq, nelmt, [x.data_ptr() for x in d_b], d_in.data_ptr(), d_out.data_ptr(),
                       [w1.data_ptr(), w2.data_ptr()], stream=st)
    got = G.host(d_out)
    want = oracle.bwdtrans_hex(nq[0], nq[1], nq[2], nelmt, b[0], b[1], b[2], inp, use_fma=True)
    assert np.array_equal(got, want), G.rel_max(got, want)
    if os.path.exists(BLAS_SO) and nq[0] == nq[1] == nq[2]:
        blas = ctypes.CDLL(BLAS_SO)
        c_out = torch.full_like(d_out, float("nan"))
        c1 = torch.empty(nelmt * nq[2] * nm[0] * nm[1], dtype=d_in.dtype, device="cuda")
        c2 = torch.empty(nelmt * nq[1] * nq[2] * nm[0], dtype=d_in.dtype, device="cuda")
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        rc = getattr(blas, f"ref_cublas_bwdtrans_hex_{suf}")(
            ctypes.c_int(nq[0]), ctypes.c_int(nq[1]), ctypes.c_int(nq[2]), ctypes.c_int(nelmt), vp(d_b[0]), vp(d_b[1]),
            vp(d_b[2]), vp(d_in), vp(c1), vp(c2), vp(c_out), ctypes.c_void_p(st))
        assert rc == 0
        assert G.rel_max(G.host(c_out), want) < (1e-12 if suf == "f64" else 1e-5)


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("M,N,batch,shared_matrix", [(4, 4, 1000, False), (9, 16, 4097, False), (31, 32, 2500, False),
                                                      (64, 48, 300, False), (1, 7, 33, False), (12, 12, 5000, True),
                                                      (128, 100, 17, False)])
def test_batched_small_matvec(G, suf, M, N, batch, shared_matrix):
    import torch
    dt = G.NP[suf]
    rng = np.random.default_rng(M * 131 + N)
    A = rng.standard_normal((1 if shared_matrix else batch, M, N)).astype(dt)
    x = rng.standard_normal((batch, N)).astype(dt)
    d_A, d_x = G.dev(A.ravel()), G.dev(x.ravel())
    st = torch.cuda.current_stream().cuda_stream
    ys = []
    for _ in range(2):
        d_y = torch.full((batch * M,), float("nan"), dtype=d_x.dtype, device="cuda")
        G.fe.matvec_batched(suf, M, N, batch, d_A.data_ptr(), 0 if shared_matrix else M * N, d_x.data_ptr(), N,
                            d_y.data_ptr(), M, stream=st)
        ys.append(G.host(d_y).reshape(batch, M))
    assert np.array_equal(ys[0], ys[1])                                     # deterministic
    Ab = np.broadcast_to(A, (batch, M, N)).astype(np.longdouble)
    want = np.einsum("bij,bj->bi", Ab, x.astype(np.longdouble))
    scale = np.einsum("bij,bj->bi", np.abs(Ab), np.abs(x).astype(np.longdouble))
    tol = 1e-12 if suf == "f64" else 1e-5
    assert float((np.abs(ys[0] - want) / scale).max()) < tol
    # one entry against the oracle's own row dots (the restatement of benchmark03's operator)
    k = batch // 2
    one = oracle.matvec(N, M, np.ascontiguousarray(Ab[k].astype(dt)).ravel(), x[k])
    assert np.abs(ys[0][k] - one).max() <= tol * float(scale[k].max())


def test_batched_matvec_argument_errors(G):
    import torch
    t = torch.zeros(1 << 20, dtype=torch.float64, device="cuda")
    E = G.fe.B200feError
    with pytest.raises(E) as e:
        G.fe.matvec_batched("f64", 0, 4, 10, t.data_ptr(), 0, t.data_ptr(), 4, t.data_ptr(), 4)
    assert e.value.code == G.fe.E_INVAL
    with pytest.raises(E) as e:                       # 400 x 400 doubles do not fit a CTA's shared memory
        G.fe.matvec_batched("f64", 400, 400, 2, t.data_ptr(), 160000, t.data_ptr(), 400, t.data_ptr(), 400)
    assert e.value.code == G.fe.E_UNSUPPORTED
    G.fe.matvec_batched("f64", 4, 4, 0, t.data_ptr(), 16, t.data_ptr(), 4, t.data_ptr(), 4)   # empty batch: no-op

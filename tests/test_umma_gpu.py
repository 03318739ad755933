"""The tcgen05 back-end ("umma", csrc/sumfac_umma.cuh): quad BwdTrans FP32 at nq = 32 on the 5th-generation tensor cores
(tcgen05.mma kind::tf32, operands / accumulators in TMEM, 3xTF32 split) -- the default route for that case.

Parity bar: north_star's FP32 tolerance, stated component-wise (include/b200fe.h):
    |out - exact| <= 1e-5 * (|B1|^T |B0|^T |in|)    for every output,
against the oracle evaluated in double (the restatement of benchmark04.cc:49-72).  Shapes: whole tiles of 4 elements,
ragged tails (the last tile is read from global memory instead of the bulk-copied slab), one element, many tiles per
CTA; a misaligned input must leave the route (the bulk copy needs 16-byte alignment) and stay correct.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
NQ, NM = 32, 31


@pytest.fixture(scope="module")
def G():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests import gpu_util
    assert gpu_util.fe.check_device() == 0, "not an sm_100 device"
    return gpu_util


def componentwise(got, b0, b1, inp, nelmt):
    want = oracle.bwdtrans_quad(NQ, NQ, nelmt, b0.astype(np.float64), b1.astype(np.float64), inp.astype(np.float64))
    scale = oracle.bwdtrans_quad(NQ, NQ, nelmt, np.abs(b0).astype(np.float64), np.abs(b1).astype(np.float64),
                                 np.abs(inp).astype(np.float64))
    assert not np.isnan(got).any()
    return float((np.abs(got.astype(np.float64) - want) / scale).max())


@pytest.mark.parametrize("nelmt", [1, 3, 4, 5, 8, 601, 4096, 4097, 148 * 4 * 3 + 2])
def test_default_route_is_umma_and_meets_the_componentwise_bound(G, nelmt):
    rng = np.random.default_rng(3200 + nelmt)
    b0 = rng.standard_normal(NM * NQ).astype(np.float32)
    b1 = rng.standard_normal(NM * NQ).astype(np.float32)
    inp = rng.standard_normal(nelmt * NM * NM).astype(np.float32)
    got = G.run_quad("BwdTransQuadKernel_QP_Shared", "f32", NQ, NQ, nelmt, b0, b1, inp)
    assert G.fe.last_backend() == "umma"
    assert componentwise(got, b0, b1, inp, nelmt) < 1e-5
    # every element-major entry point takes the same route
    for k in ("BwdTransQuadKernel", "BwdTransQuadKernel_QP", "BwdTransQuadKernel_QP_1D", "BwdTransQuadKernel_QP_1D_Shared"):
        again = G.run_quad(k, "f32", NQ, NQ, nelmt, b0, b1, inp)
        assert G.fe.last_backend() == "umma" and np.array_equal(again, got)     # deterministic, same kernel


def test_reference_synthetic_input_and_golden_norm(G, golden):
    """the reference's own data (sin / cos tables, heavy cancellation): FP32 norm against the FP64 golden value"""
    nelmt = 4096
    b = oracle.gen_basis(NM, NQ, np.float32)
    inp = oracle.gen_in(nelmt, NM * NM, np.float32)
    got = G.run_quad("BwdTransQuadKernel_QP_Shared", "f32", NQ, NQ, nelmt, b, b, inp)
    assert G.fe.last_backend() == "umma"
    assert componentwise(got, b, b, inp, nelmt) < 1e-5
    norm = float(np.sqrt(oracle.sumsq(got.astype(np.float64))))
    want = golden["quad"]["32"][str(nelmt)][0]
    assert abs(norm - want) / want < 1e-5


def test_misaligned_input_leaves_the_route_and_stays_correct(G):
    import torch
    nelmt = 64
    rng = np.random.default_rng(77)
    b0 = rng.standard_normal(NM * NQ).astype(np.float32)
    b1 = rng.standard_normal(NM * NQ).astype(np.float32)
    inp = rng.standard_normal(nelmt * NM * NM).astype(np.float32)
    big = torch.zeros(inp.size + 4, dtype=torch.float32, device="cuda")
    big[1:1 + inp.size] = torch.from_numpy(inp).cuda()
    d_b0, d_b1 = G.dev(b0), G.dev(b1)
    d_out = torch.full((nelmt * NQ * NQ,), float("nan"), dtype=torch.float32, device="cuda")
    G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f32", NQ, NQ, nelmt, d_b0.data_ptr(), d_b1.data_ptr(),
                       big.data_ptr() + 4, d_out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    assert G.fe.last_backend() != "umma"
    assert componentwise(G.host(d_out), b0, b1, inp, nelmt) < 1e-5
    with pytest.raises(G.fe.B200feError) as e:          # forced: refused, not silently rerouted
        G.fe.set_backend("umma")
        try:
            G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f32", NQ, NQ, nelmt, d_b0.data_ptr(), d_b1.data_ptr(),
                               big.data_ptr() + 4, d_out.data_ptr())
        finally:
            G.fe.set_backend("auto")
    assert e.value.code == G.fe.E_UNSUPPORTED


def test_forced_on_other_shapes_is_refused(G):
    import torch
    t = torch.zeros(1 << 16, dtype=torch.float32, device="cuda")
    d = torch.zeros(1 << 16, dtype=torch.float64, device="cuda")
    G.fe.set_backend("umma")
    try:
        for call in (lambda: G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f32", 16, 16, 8, t.data_ptr(), t.data_ptr(),
                                                t.data_ptr(), t.data_ptr()),
                     lambda: G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f64", 32, 32, 4, d.data_ptr(), d.data_ptr(),
                                                d.data_ptr(), d.data_ptr()),
                     lambda: G.fe.bwdtrans_quad("BwdTransQuadKernel_Coa", "f32", 32, 32, 32, t.data_ptr(), t.data_ptr(),
                                                t.data_ptr(), t.data_ptr()),
                     lambda: G.fe.bwdtrans_hex("BwdTransHexKernel_QP_Shared", "f32", 8, 8, 8, 4, t.data_ptr(), t.data_ptr(),
                                               t.data_ptr(), t.data_ptr(), t.data_ptr())):
            with pytest.raises(G.fe.B200feError) as e:
                call()
            assert e.value.code == G.fe.E_UNSUPPORTED
    finally:
        G.fe.set_backend("auto")


def test_fused_checksum_and_plan_take_the_same_kernel(G):
    import torch
    nelmt = 2050
    rng = np.random.default_rng(5)
    b0 = rng.standard_normal(NM * NQ).astype(np.float32)
    b1 = rng.standard_normal(NM * NQ).astype(np.float32)
    inp = rng.standard_normal(nelmt * NM * NM).astype(np.float32)
    plain = G.run_quad("BwdTransQuadKernel_QP_Shared", "f32", NQ, NQ, nelmt, b0, b1, inp)
    st = torch.cuda.current_stream().cuda_stream
    d_b0, d_b1, d_in = G.dev(b0), G.dev(b1), G.dev(inp)
    d_out = torch.full((nelmt * NQ * NQ,), float("nan"), dtype=torch.float32, device="cuda")
    ss = torch.full((1,), float("nan"), dtype=torch.float64, device="cuda")
    scratch = torch.empty(G.fe.sumsq_scratch_bytes(), dtype=torch.uint8, device="cuda")
    G.fe.bwdtrans_sumsq("f32", (NQ, NQ), nelmt, [d_b0.data_ptr(), d_b1.data_ptr()], d_in.data_ptr(), d_out.data_ptr(),
                        ss.data_ptr(), scratch.data_ptr(), st)
    fused = G.host(d_out)
    assert G.fe.last_backend() == "umma" and np.array_equal(fused, plain)
    want = oracle.sumsq(fused)
    assert abs(float(ss.item()) - want) / want < 1e-12
    plan = G.fe.Plan(2, "f32", NQ, [d_b0.data_ptr(), d_b1.data_ptr()], stream=st)
    d_out.fill_(float("nan"))
    plan.bwdtrans(nelmt, d_in.data_ptr(), d_out.data_ptr(), stream=st)
    assert G.fe.last_backend() == "umma" and np.array_equal(G.host(d_out), plain)
    plan.destroy()


def test_back_to_back_launches_with_different_data_do_not_interfere(G):
    """TMEM is allocated and released by every CTA; 50 launches alternating two inputs and two bases on two streams"""
    import torch
    nelmt = 1184
    rng = np.random.default_rng(9)
    sets = []
    for _ in range(2):
        b0 = rng.standard_normal(NM * NQ).astype(np.float32)
        b1 = rng.standard_normal(NM * NQ).astype(np.float32)
        inp = rng.standard_normal(nelmt * NM * NM).astype(np.float32)
        sets.append((G.dev(b0), G.dev(b1), G.dev(inp), G.run_quad("BwdTransQuadKernel_QP_Shared", "f32", NQ, NQ, nelmt, b0, b1, inp)))
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [torch.empty(nelmt * NQ * NQ, dtype=torch.float32, device="cuda") for _ in range(50)]
    torch.cuda.synchronize()
    for k in range(50):
        d_b0, d_b1, d_in, _ = sets[k % 2]
        G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f32", NQ, NQ, nelmt, d_b0.data_ptr(), d_b1.data_ptr(),
                           d_in.data_ptr(), outs[k].data_ptr(), stream=streams[k % 3 % 2].cuda_stream)
    torch.cuda.synchronize()
    for k in range(50):
        assert np.array_equal(outs[k].cpu().numpy(), sets[k % 2][3]), k

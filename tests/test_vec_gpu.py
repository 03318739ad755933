"""Parity of the benchmark01-03 kernels, the checksum and the host-buffer
pipeline with the CPU oracle, through the C ABI."""
import math

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests import gpu_util
    return gpu_util


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("n", [1, 3, 1024, 100191 * 2 + 5, 1 << 20])
def test_set_data_bit_exact(G, suf, n):
    import torch
    dt = G.NP[suf]
    for second, hostgen in ((False, False), (False, True), (True, False)):
        d = torch.empty(n, dtype=torch.float64 if suf == "f64" else torch.float32, device="cuda")
        G.fe.set_data(suf, d.data_ptr(), n, second=second, hostgen=hostgen, stream=stream())
        # b200fe_set_data == the reference's device kernel (fused), _hostgen/_set_data2 == benchmark02's host loops
        want = oracle.set_data(n, dt, second=second, fused=not (second or hostgen))
        assert np.array_equal(G.host(d), want)


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("n", [1, 2, 3, 31, 1023, 1024, 1025, 65537, (1 << 22) + 3])
def test_l2norm_all_variants(G, suf, n):
    import torch
    dt = G.NP[suf]
    tdt = torch.float64 if suf == "f64" else torch.float32
    x = oracle.set_data(n, dt)
    want = oracle.sumsq(x)
    d = G.dev(x)
    blocks = min((n + 255) // 256, 1024)
    tol = 1e-12 if suf == "f64" else 1e-5
    for variant in ("scalar", "vl", "functor"):
        sums = torch.full((blocks,), float("nan"), dtype=tdt, device="cuda")
        res = torch.full((1,), float("nan"), dtype=tdt, device="cuda")
        if variant == "functor":
            G.fe.reduce_sum_sumsq(suf, 0, n, sums.data_ptr(), d.data_ptr(), blocks, stream())
        else:
            G.fe.l2norm_vl(suf, sums.data_ptr(), d.data_ptr(), n, blocks, variant == "vl", stream())
        G.fe.reduce_vl(suf, res.data_ptr(), sums.data_ptr(), blocks, variant != "scalar", stream())
        got = float(res.item())
        assert abs(got - want) / want < tol, (variant, n, got, want)
        # deterministic: a second run gives the same bits (the reference's atomics do not)
        sums2 = torch.empty_like(sums)
        if variant == "functor":
            G.fe.reduce_sum_sumsq(suf, 0, n, sums2.data_ptr(), d.data_ptr(), blocks, stream())
        else:
            G.fe.l2norm_vl(suf, sums2.data_ptr(), d.data_ptr(), n, blocks, variant == "vl", stream())
        assert torch.equal(sums, sums2)


def test_b01_golden_norm_and_subrange(G, golden):
    import torch
    n = 1 << 24
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    G.fe.set_data("f64", d.data_ptr(), n, stream=stream())
    sums = torch.empty(1024, dtype=torch.float64, device="cuda")
    res = torch.empty(1, dtype=torch.float64, device="cuda")
    G.fe.l2norm_vl("f64", sums.data_ptr(), d.data_ptr(), n, 1024, True, stream())
    G.fe.reduce_vl("f64", res.data_ptr(), sums.data_ptr(), 1024, True, stream())
    for want in golden["b01"][str(n)]:
        assert abs(math.sqrt(float(res.item())) - want) / want < 6e-10
    # [begin, end) with an odd begin: the functor kernel's sub-range (multi-GPU sharding uses it)
    b, e = 12345, 7654321
    G.fe.reduce_sum_sumsq("f64", b, e, sums.data_ptr(), d.data_ptr(), 1024, stream())
    G.fe.reduce_vl("f64", res.data_ptr(), sums.data_ptr(), 1024, True, stream())
    want = oracle.sumsq(oracle.set_data(n, fused=True)[b:e])
    assert abs(float(res.item()) - want) / want < 1e-12


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("n", [1, 5, 1024, 4099, (1 << 21) + 1])
def test_add_vector_bit_exact_40_reps(G, suf, n):
    dt = G.NP[suf]
    x0, y = oracle.set_data(n, dt), oracle.set_data(n, dt, second=True)
    want = oracle.add_vector(x0.copy(), y, reps=40)
    for variant in ("scalar", "vl", "functor"):
        dx, dy = G.dev(x0), G.dev(y)
        for _ in range(40):
            if variant == "functor":
                G.fe.vector_kernel_add(suf, 0, n, dx.data_ptr(), dy.data_ptr(), stream())
            else:
                G.fe.add_vector(suf, dx.data_ptr(), dy.data_ptr(), n, variant == "vl", stream())
        assert np.array_equal(G.host(dx), want), (variant, n)


def test_b02_golden_norm(G, golden):
    import torch
    n = 1 << 22
    dx = torch.empty(n, dtype=torch.float64, device="cuda")
    dy = torch.empty(n, dtype=torch.float64, device="cuda")
    G.fe.set_data("f64", dx.data_ptr(), n, hostgen=True, stream=stream())  # benchmark02 initialises on the host
    G.fe.set_data("f64", dy.data_ptr(), n, second=True, stream=stream())
    for _ in range(40):
        G.fe.add_vector("f64", dx.data_ptr(), dy.data_ptr(), n, True, stream())
    res = torch.empty(1, dtype=torch.float64, device="cuda")
    scratch = torch.empty(G.fe.sumsq_scratch_bytes(), dtype=torch.uint8, device="cuda")
    G.fe.sumsq("f64", dx.data_ptr(), n, res.data_ptr(), scratch.data_ptr(), stream())
    for want in golden["b02"][str(n)]:
        assert abs(math.sqrt(float(res.item())) - want) / want < 6e-10


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("shape", [(128, 128), (200, 77), (1, 5000), (37, 4096), (16, 5001)])
def test_matvec(G, suf, shape):
    import torch
    M, N = shape
    dt = G.NP[suf]
    rng = np.random.default_rng(M * 131 + N)
    A, x = rng.standard_normal(M * N).astype(dt), rng.standard_normal(N).astype(dt)
    want = oracle.matvec(N, M, A, x)
    scale = (np.abs(A.reshape(M, N).astype(np.float64)) @ np.abs(x.astype(np.float64)))
    tol = 1e-12 if suf == "f64" else 1e-5
    for vl in (False, True):
        dA, dx = G.dev(A), G.dev(x)
        dy = torch.full((M,), float("nan"), dtype=dA.dtype, device="cuda")
        G.fe.compute_matvec(suf, N, M, dA.data_ptr(), dx.data_ptr(), dy.data_ptr(), vl, stream())
        got = G.host(dy)
        assert np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= tol * scale), (shape, vl)


@pytest.mark.parametrize("n", [128, 1024, 4096])
def test_b03_golden_norm(G, golden, n):
    import torch
    A, x = oracle.gen_matvec(n, n)
    dA, dx = G.dev(A), G.dev(x)
    dy = torch.empty(n, dtype=torch.float64, device="cuda")
    G.fe.compute_matvec("f64", n, n, dA.data_ptr(), dx.data_ptr(), dy.data_ptr(), True, stream())
    got = math.sqrt(oracle.sumsq(G.host(dy)))
    for want in golden["b03"][str(n)]:
        assert abs(got - want) / want < 5e-9


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("n", [1, 7, 4097, (1 << 23) + 11])
def test_sumsq_checksum(G, suf, n):
    import torch
    dt = G.NP[suf]
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(dt)
    d = G.dev(x)
    res = torch.full((1,), float("nan"), dtype=torch.float64, device="cuda")
    scratch = torch.empty(G.fe.sumsq_scratch_bytes(), dtype=torch.uint8, device="cuda")
    G.fe.sumsq(suf, d.data_ptr(), n, res.data_ptr(), scratch.data_ptr(), stream())
    want = oracle.sumsq(x)
    assert abs(float(res.item()) - want) / want < 1e-12


@pytest.mark.parametrize("suf", ["f64", "f32"])
def test_host_buffer_pipeline_matches_oracle(G, suf):
    import torch
    dt = G.NP[suf]
    rng = np.random.default_rng(21)
    # hex nq=4: chunks of (24 MiB / 27 values) elements -> several chunks + ragged tail
    nq, nm = 4, 3
    nelmt = 300000 if suf == "f64" else 500000
    b = [rng.standard_normal(nm * nq).astype(dt) for _ in range(3)]
    tdt = torch.float64 if suf == "f64" else torch.float32
    h_in = torch.from_numpy(rng.standard_normal(nelmt * nm ** 3).astype(dt)).pin_memory()
    h_out = torch.empty(nelmt * nq ** 3, dtype=tdt).pin_memory()
    hb = [torch.from_numpy(x) for x in b]
    ss = G.fe.bwdtrans_host(suf, (nq, nq, nq), nelmt, [t.data_ptr() for t in hb], h_in.data_ptr(), h_out.data_ptr())
    want = oracle.bwdtrans_hex(nq, nq, nq, nelmt, *b, h_in.numpy())
    assert np.array_equal(h_out.numpy(), want)
    assert abs(ss - oracle.sumsq(want)) / oracle.sumsq(want) < 1e-12
    # checksum-only form (no copy-back), quad
    nq = 6
    nm = 5
    nelmt = 200001
    b2 = [rng.standard_normal(nm * nq).astype(dt) for _ in range(2)]
    h_in2 = torch.from_numpy(rng.standard_normal(nelmt * nm * nm).astype(dt)).pin_memory()
    hb2 = [torch.from_numpy(x) for x in b2]
    ss2 = G.fe.bwdtrans_host(suf, (nq, nq), nelmt, [t.data_ptr() for t in hb2], h_in2.data_ptr(), 0)
    want2 = oracle.sumsq(oracle.bwdtrans_quad(nq, nq, nelmt, *b2, h_in2.numpy()))
    assert abs(ss2 - want2) / want2 < (1e-12 if suf == "f64" else 1e-6)

"""The interleaved ("Coales") entry points at the largest nq of the BASELINE sweeps, where round 2 added two
back-ends (reference layout and kernels: benchmark04.cc:78-147, benchmark05.cc:104-201 with the offsets of :810-812):

  coa-pipe  FP64 hexes nq = 8, 10 (csrc/sumfac_coapipe.cuh): persistent CTAs, cp.async gather of [idx][e] tiles, three
            passes that keep the interleaved order, pass 1 in place          -- b200fe_set_backend("pipe")
  coa-mma   quads nq = 32 (csrc/sumfac_coamma.cuh): tensor cores with M = elements, t1 in place -- b200fe_set_backend("mma")
            FP64 on DMMA; FP32 on mma.sync with the 3xTF32 split

The FP64 kernels accumulate in the reference's order with fused multiply-adds and must reproduce the oracle BIT FOR
BIT; the FP32 tensor-core kernel is held to the component-wise bound of include/b200fe.h,
|out - exact| <= 1e-5 * (|B1|^T |B0|^T |in|) for every single output (exact = the oracle in double).  Sizes
are chosen so that the persistent CTAs loop over several tiles (more tiles than resident CTAs), inputs differ per
element and interleave group, and an element with non-finite modes must not leak into its neighbours (the tensor-core
kernel pads K with a value it reads from shared memory and zero-selects).
"""
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
    assert gpu_util.fe.check_device() == 0, "not an sm_100 device"
    return gpu_util


def rnd(rng, n, dt):
    return rng.standard_normal(n).astype(dt)


def quad_case(nq, nelmt, seed):
    nm, rng = nq - 1, np.random.default_rng(seed)
    b0, b1 = rnd(rng, nm * nq, np.float64), rnd(rng, nm * nq, np.float64)
    inp_em = rnd(rng, nelmt * nm * nm, np.float64)
    return b0, b1, inp_em


def hex_case(nq, nelmt, seed):
    nm, rng = nq - 1, np.random.default_rng(seed)
    b = [rnd(rng, nm * nq, np.float64) for _ in range(3)]
    inp_em = rnd(rng, nelmt * nm ** 3, np.float64)
    return b, inp_em


@pytest.mark.parametrize("nelmt", [32, 32 * 7, 32 * 151])  # 151 groups = 604 tiles of 8 > 2 x 148 resident CTAs
def test_quad_nq32_fp64_tensor_core_kernel_bit_exact(G, nelmt):
    nq = 32
    b0, b1, inp_em = quad_case(nq, nelmt, 4100 + nelmt)
    want_em = oracle.bwdtrans_quad(nq, nq, nelmt, b0, b1, inp_em)
    inp = oracle.to_coa(inp_em, nelmt, (nq - 1) ** 2)
    got = G.run_quad("BwdTransQuadKernel_Coa", "f64", nq, nq, nelmt, b0, b1, inp)
    assert G.fe.last_backend() == "coa-mma"  # the default route
    assert np.array_equal(oracle.from_coa(got, nelmt, nq * nq), want_em)
    try:
        G.fe.set_backend("mma")
        got = G.run_quad("BwdTransQuadKernel_Coa", "f64", nq, nq, nelmt, b0, b1, inp)
        assert G.fe.last_backend() == "coa-mma"
        assert np.array_equal(oracle.from_coa(got, nelmt, nq * nq), want_em)
        G.fe.set_backend("lanes")  # the kernel it replaced stays available and agrees bit for bit
        got = G.run_quad("BwdTransQuadKernel_Coa", "f64", nq, nq, nelmt, b0, b1, inp)
        assert G.fe.last_backend() == "lanes"
        assert np.array_equal(oracle.from_coa(got, nelmt, nq * nq), want_em)
    finally:
        G.fe.set_backend("auto")


@pytest.mark.parametrize("nq,nelmt", [(10, 32), (10, 32 * 5), (10, 32 * 75), (8, 32), (8, 32 * 9), (8, 32 * 240)])
def test_hex_fp64_coa_pipe_kernel_bit_exact(G, nq, nelmt):
    # nq = 10: tiles of 16, one CTA per SM: 150 tiles > 148;  nq = 8: tiles of 8, <= 4 CTAs per SM: 960 tiles > 592
    b, inp_em = hex_case(nq, nelmt, 4200 + nq + nelmt)
    want_em = oracle.bwdtrans_hex(nq, nq, nq, nelmt, *b, inp_em)
    inp = oracle.to_coa(inp_em, nelmt, (nq - 1) ** 3)
    got = G.run_hex("BwdTransHexKernel_Coa", "f64", (nq, nq, nq), nelmt, b, inp)
    assert G.fe.last_backend() == "coa-pipe"
    assert np.array_equal(oracle.from_coa(got, nelmt, nq ** 3), want_em)
    try:
        G.fe.set_backend("pipe")
        got = G.run_hex("BwdTransHexKernel_Coa", "f64", (nq, nq, nq), nelmt, b, inp)
        assert G.fe.last_backend() == "coa-pipe"
        assert np.array_equal(oracle.from_coa(got, nelmt, nq ** 3), want_em)
        G.fe.set_backend("lanes")
        got = G.run_hex("BwdTransHexKernel_Coa", "f64", (nq, nq, nq), nelmt, b, inp)
        assert G.fe.last_backend() == "lanes"
        assert np.array_equal(oracle.from_coa(got, nelmt, nq ** 3), want_em)
    finally:
        G.fe.set_backend("auto")


def componentwise_quad(got_em, nq, nelmt, b0, b1, inp_em):
    d = np.float64
    want = oracle.bwdtrans_quad(nq, nq, nelmt, b0.astype(d), b1.astype(d), inp_em.astype(d))
    scale = oracle.bwdtrans_quad(nq, nq, nelmt, np.abs(b0).astype(d), np.abs(b1).astype(d), np.abs(inp_em).astype(d))
    assert not np.isnan(got_em).any()
    return float((np.abs(got_em.astype(d) - want) / scale).max())


@pytest.mark.parametrize("nelmt", [32, 32 * 7, 32 * 151])  # 302 tiles of 16 > 2 x 148 resident CTAs
def test_quad_nq32_fp32_tensor_core_kernel_meets_the_componentwise_bound(G, nelmt):
    nq = 32
    b0, b1, inp_em = [x.astype(np.float32) for x in quad_case(nq, nelmt, 4150 + nelmt)]
    inp = oracle.to_coa(inp_em, nelmt, (nq - 1) ** 2)
    got = G.run_quad("BwdTransQuadKernel_Coa", "f32", nq, nq, nelmt, b0, b1, inp)
    assert G.fe.last_backend() == "coa-mma"  # the default route
    err = componentwise_quad(oracle.from_coa(got, nelmt, nq * nq), nq, nelmt, b0, b1, inp_em)
    assert err < 1e-5, err
    again = G.run_quad("BwdTransQuadKernel_Coa", "f32", nq, nq, nelmt, b0, b1, inp)
    assert np.array_equal(again, got)  # deterministic
    try:
        G.fe.set_backend("lanes")  # the bit-exact FP32 kernel stays available
        exact = G.run_quad("BwdTransQuadKernel_Coa", "f32", nq, nq, nelmt, b0, b1, inp)
        assert G.fe.last_backend() == "lanes"
        assert np.array_equal(oracle.from_coa(exact, nelmt, nq * nq), oracle.bwdtrans_quad(nq, nq, nelmt, b0, b1, inp_em))
    finally:
        G.fe.set_backend("auto")


def test_quad_nq32_fp32_reference_synthetic_input(G):
    """the reference's own synthetic data (benchmark04.cc:831-850: sin / cos tables, heavy cancellation) through the
    oracle's generators: still inside the component-wise bound"""
    nq, nelmt = 32, 64
    b0 = oracle.gen_basis(nq - 1, nq, np.float32)
    inp_em = oracle.gen_in(nelmt, (nq - 1) ** 2, np.float32)
    got = G.run_quad("BwdTransQuadKernel_Coa", "f32", nq, nq, nelmt, b0, b0, oracle.to_coa(inp_em, nelmt, (nq - 1) ** 2))
    assert G.fe.last_backend() == "coa-mma"
    assert componentwise_quad(oracle.from_coa(got, nelmt, nq * nq), nq, nelmt, b0, b0, inp_em) < 1e-5


@pytest.mark.parametrize("nq,nelmt", [(10, 32 * 80), (8, 32 * 250)])
def test_hex_coa_pipe_gather_routes_agree(G, nq, nelmt):
    """the tile is gathered by tiled TMA through a tensor map (default) or, where the driver does not offer the
    encoder, by 16-byte cp.async copies: same kernel otherwise, same bits -- b200fe_set_gather forces either"""
    assert G.fe.tensor_map_available(), "the driver on a B200 box offers cuTensorMapEncodeTiled: the TMA route must be live"
    b, inp_em = hex_case(nq, nelmt, 4250 + nq)
    want_em = oracle.bwdtrans_hex(nq, nq, nq, nelmt, *b, inp_em)
    inp = oracle.to_coa(inp_em, nelmt, (nq - 1) ** 3)
    try:
        for mode in ("cp.async", "tma"):
            G.fe.set_gather(mode)
            n0 = G.fe.launch_count()
            got = G.run_hex("BwdTransHexKernel_Coa", "f64", (nq, nq, nq), nelmt, b, inp)
            assert G.fe.last_backend() == "coa-pipe" and G.fe.launch_count() - n0 == 2  # bank fill + operator
            assert np.array_equal(oracle.from_coa(got, nelmt, nq ** 3), want_em), mode
        with pytest.raises(Exception):
            G.fe.set_gather("ldg")
    finally:
        G.fe.set_gather("tma")


def test_forced_back_ends_where_they_have_no_instantiation(G):
    b0, b1, inp_em = quad_case(16, 32, 1)
    bh, inph = hex_case(6, 32, 2)
    try:
        G.fe.set_backend("mma")
        with pytest.raises(Exception):  # interleaved quads on the tensor cores: nq = 32 only
            G.run_quad("BwdTransQuadKernel_Coa", "f64", 16, 16, 32, b0, b1, oracle.to_coa(inp_em, 32, 225))
        G.fe.set_backend("pipe")
        with pytest.raises(Exception):  # interleaved hexes: FP64 nq = 8, 10 only
            G.run_hex("BwdTransHexKernel_Coa", "f64", (6, 6, 6), 32, bh, oracle.to_coa(inph, 32, 125))
        with pytest.raises(Exception):
            G.run_hex("BwdTransHexKernel_Coa", "f32", (10, 10, 10), 32, *[[x.astype(np.float32) for x in hex_case(10, 32, 4)[0]],
                      hex_case(10, 32, 4)[1].astype(np.float32)])
        with pytest.raises(Exception):  # quads have no interleaved pipe kernel
            G.run_quad("BwdTransQuadKernel_Coa", "f64", 16, 16, 32, b0, b1, oracle.to_coa(inp_em, 32, 225))
    finally:
        G.fe.set_backend("auto")


@pytest.mark.parametrize("which", ["quad32", "hex10", "hex8"])
def test_non_finite_elements_stay_isolated(G, which):
    """one element full of Inf / NaN: every other element of its tile, group and neighbourhood is still bit-exact"""
    nelmt, bad = 32 * 3, 41  # element 41 = group 1, lane 9: second tile of 8, first of 16
    if which == "quad32":
        nq = 32
        b0, b1, inp_em = quad_case(nq, nelmt, 4300)
        per = (nq - 1) ** 2
        inp_em.reshape(nelmt, per)[bad, ::2] = np.inf
        inp_em.reshape(nelmt, per)[bad, 1::2] = np.nan
        want = oracle.bwdtrans_quad(nq, nq, nelmt, b0, b1, inp_em).reshape(nelmt, -1)
        got = G.run_quad("BwdTransQuadKernel_Coa", "f64", nq, nq, nelmt, b0, b1, oracle.to_coa(inp_em, nelmt, per))
        assert G.fe.last_backend() == "coa-mma"
        got = oracle.from_coa(got, nelmt, nq * nq).reshape(nelmt, -1)
    else:
        nq = 10 if which == "hex10" else 8
        b, inp_em = hex_case(nq, nelmt, 4301)
        per = (nq - 1) ** 3
        inp_em.reshape(nelmt, per)[bad, ::2] = np.inf
        inp_em.reshape(nelmt, per)[bad, 1::2] = np.nan
        want = oracle.bwdtrans_hex(nq, nq, nq, nelmt, *b, inp_em).reshape(nelmt, -1)
        got = G.run_hex("BwdTransHexKernel_Coa", "f64", (nq, nq, nq), nelmt, b, oracle.to_coa(inp_em, nelmt, per))
        assert G.fe.last_backend() == "coa-pipe"
        got = oracle.from_coa(got, nelmt, nq ** 3).reshape(nelmt, -1)
    ok = np.ones(nelmt, bool)
    ok[bad] = False
    assert np.isfinite(got[ok]).all()
    assert np.array_equal(got[ok], want[ok])
    assert not np.isfinite(got[bad]).any()


@pytest.mark.parametrize("which", ["quad32", "hex10"])
def test_misaligned_input_falls_back_and_nothing_is_written_outside_out(G, which):
    """the gather moves 16-byte chunks: an `in` that is only 8-byte aligned takes the lanes kernel under the default
    routing and is refused when the new back-end is forced; `out` may be misaligned (8-byte stores)"""
    import torch
    nelmt = 64
    st = torch.cuda.current_stream().cuda_stream
    if which == "quad32":
        nq, dim = 32, 2
        b0, b1, inp_em = quad_case(nq, nelmt, 4400)
        bs = [b0, b1]
        want = oracle.to_coa(oracle.bwdtrans_quad(nq, nq, nelmt, b0, b1, inp_em), nelmt, nq * nq)
    else:
        nq, dim = 10, 3
        bs, inp_em = hex_case(nq, nelmt, 4401)
        want = oracle.to_coa(oracle.bwdtrans_hex(nq, nq, nq, nelmt, *bs, inp_em), nelmt, nq ** 3)
    inp = oracle.to_coa(inp_em, nelmt, (nq - 1) ** dim)
    d_b = [G.dev(x) for x in bs]
    big_in = torch.zeros(inp.size + 4, dtype=torch.float64, device="cuda")
    big_in[1:1 + inp.size] = G.dev(inp)
    d_out = torch.full((want.size + 2,), float("nan"), dtype=torch.float64, device="cuda")

    def call(in_ptr, out_ptr):
        if dim == 2:
            G.fe.bwdtrans_quad("BwdTransQuadKernel_Coa", "f64", nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(), in_ptr,
                               out_ptr, stream=st)
        else:
            G.fe.bwdtrans_hex("BwdTransHexKernel_Coa", "f64", nq, nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                              d_b[2].data_ptr(), in_ptr, out_ptr, stream=st)

    call(big_in.data_ptr() + 8, d_out.data_ptr() + 8)
    assert G.fe.last_backend() == "lanes"
    got = G.host(d_out)
    assert np.array_equal(got[1:-1], want) and np.isnan(got[0]) and np.isnan(got[-1])
    try:
        G.fe.set_backend("mma" if dim == 2 else "pipe")
        with pytest.raises(Exception):
            call(big_in.data_ptr() + 8, d_out.data_ptr() + 8)
    finally:
        G.fe.set_backend("auto")
    d_in = G.dev(inp)
    d_out.fill_(float("nan"))
    call(d_in.data_ptr(), d_out.data_ptr() + 8)
    assert G.fe.last_backend() == ("coa-mma" if dim == 2 else "coa-pipe")
    got = G.host(d_out)
    assert np.array_equal(got[1:-1], want) and np.isnan(got[0]) and np.isnan(got[-1])


@pytest.mark.parametrize("which", ["quad32", "hex10"])
def test_baseline_size_every_element_equal(G, which):
    """64 Mi points through a size-independent property: every element gets the same modes, so every (group, lane)
    must reproduce element 0 of the oracle exactly"""
    import torch
    dim, nq = (2, 32) if which == "quad32" else (3, 10)
    nm = nq - 1
    nelmt = (64 << 20) // nq ** dim // 32 * 32
    rng = np.random.default_rng(4500 + nq)
    b = [rnd(rng, nm * nq, np.float64) for _ in range(dim)]
    one = rnd(rng, nm ** dim, np.float64)
    d_in = G.dev(one).repeat_interleave(32).reshape(1, -1).repeat(nelmt // 32, 1).reshape(-1).contiguous()
    d_b = [G.dev(x) for x in b]
    d_out = torch.empty(nelmt * nq ** dim, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    if dim == 2:
        want = oracle.bwdtrans_quad(nq, nq, 1, b[0], b[1], one)
        G.fe.bwdtrans_quad("BwdTransQuadKernel_Coa", "f64", nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                           d_in.data_ptr(), d_out.data_ptr(), stream=st)
        assert G.fe.last_backend() == "coa-mma"
    else:
        want = oracle.bwdtrans_hex(nq, nq, nq, 1, *b, one)
        G.fe.bwdtrans_hex("BwdTransHexKernel_Coa", "f64", nq, nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                          d_b[2].data_ptr(), d_in.data_ptr(), d_out.data_ptr(), stream=st)
        assert G.fe.last_backend() == "coa-pipe"
    torch.cuda.synchronize()
    view = d_out.reshape(nelmt // 32, nq ** dim, 32)
    assert bool((view == G.dev(want).reshape(1, -1, 1)).all())


def test_quad_nq32_fp32_baseline_size_every_element_equal(G):
    """64 Mi points on the FP32 tensor-core kernel: identical elements must give identical outputs in every (group,
    lane) -- bit for bit, the kernel is deterministic -- and element 0 meets the component-wise bound"""
    import torch
    nq, nm = 32, 31
    nelmt = (64 << 20) // nq ** 2 // 32 * 32
    rng = np.random.default_rng(4600)
    b0, b1 = rnd(rng, nm * nq, np.float32), rnd(rng, nm * nq, np.float32)
    one = rnd(rng, nm * nm, np.float32)
    d_in = G.dev(one).repeat_interleave(32).reshape(1, -1).repeat(nelmt // 32, 1).reshape(-1).contiguous()
    d_b0, d_b1 = G.dev(b0), G.dev(b1)
    d_out = torch.empty(nelmt * nq * nq, dtype=torch.float32, device="cuda")
    G.fe.bwdtrans_quad("BwdTransQuadKernel_Coa", "f32", nq, nq, nelmt, d_b0.data_ptr(), d_b1.data_ptr(), d_in.data_ptr(),
                       d_out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    assert G.fe.last_backend() == "coa-mma"
    torch.cuda.synchronize()
    view = d_out.reshape(nelmt // 32, nq * nq, 32)
    first = view[0, :, 0].clone()
    assert bool((view == first.reshape(1, -1, 1)).all())
    assert componentwise_quad(G.host(first), nq, 1, b0, b1, one) < 1e-5

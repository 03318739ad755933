"""The warp-interleaved ("Coales") entry points, back-end by back-end, through the C ABI -- and the element-major
twin of the lanes kernel ("lanes-em").

Layout (reference benchmark04.cc:78-147, benchmark05.cc:104-201): x[(e/32)*32*len + 32*idx + e%32].
Three back-ends serve it -- thread per element (tpe, small nq), lanes (lanes = elements, an element's planes /
rows split over the warps of a CTA; the default from nq = 4), rows through a gather (the remaining nq) -- and
every one of them must reproduce the oracle BIT FOR BIT: they all accumulate in the reference's order with fused
multiply-adds.  Inputs differ per element and per interleave group, so a wrong lane / group offset shows.
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


def quad_case(G, suf, nq, nelmt, seed):
    dt, nm = G.NP[suf], nq - 1
    rng = np.random.default_rng(seed)
    b0, b1 = rnd(rng, nm * nq, dt), rnd(rng, nm * nq, dt)
    inp_em = rnd(rng, nelmt * nm * nm, dt)
    want_em = oracle.bwdtrans_quad(nq, nq, nelmt, b0, b1, inp_em)
    return b0, b1, oracle.to_coa(inp_em, nelmt, nm * nm), want_em


def hex_case(G, suf, nq, nelmt, seed):
    dt, nm = G.NP[suf], nq - 1
    rng = np.random.default_rng(seed)
    b = [rnd(rng, nm * nq, dt) for _ in range(3)]
    inp_em = rnd(rng, nelmt * nm ** 3, dt)
    want_em = oracle.bwdtrans_hex(nq, nq, nq, nelmt, *b, inp_em)
    return b, oracle.to_coa(inp_em, nelmt, nm ** 3), want_em


QUAD_LANES_NQ = list(range(4, 17)) + [32]
HEX_LANES_NQ = list(range(4, 11))


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("nq", QUAD_LANES_NQ)
def test_quad_lanes_backend_bit_exact(G, suf, nq):
    nelmt = 32 * 7  # several groups, not a multiple of anything else
    b0, b1, inp, want_em = quad_case(G, suf, nq, nelmt, 3000 + nq)
    try:
        G.fe.set_backend("lanes")
        got = G.run_quad("BwdTransQuadKernel_Coa", suf, nq, nq, nelmt, b0, b1, inp)
        assert G.fe.last_backend() == "lanes"
    finally:
        G.fe.set_backend("auto")
    assert np.array_equal(oracle.from_coa(got, nelmt, nq * nq), want_em)
    # the default routing: lanes, except FP32 nq <= 6 where the thread-per-element kernel measured faster and FP64
    # nq = 32 on the tensor cores (tests/test_coa_large_nq_gpu.py)
    got = G.run_quad("BwdTransQuadKernel_Coa", suf, nq, nq, nelmt, b0, b1, inp)
    assert G.fe.last_backend() == ("tpe" if suf == "f32" and nq <= 6 else "coa-mma" if nq == 32 else "lanes")
    if suf == "f32" and nq == 32:  # 3xTF32 tensor-core route: to rounding (tests/test_coa_large_nq_gpu.py holds the bound)
        G.assert_parity(oracle.from_coa(got, nelmt, nq * nq), want_em, suf)
    else:
        assert np.array_equal(oracle.from_coa(got, nelmt, nq * nq), want_em)


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("nq", HEX_LANES_NQ)
def test_hex_lanes_backend_bit_exact(G, suf, nq):
    nelmt = 32 * 5
    b, inp, want_em = hex_case(G, suf, nq, nelmt, 3100 + nq)
    try:
        G.fe.set_backend("lanes")
        got = G.run_hex("BwdTransHexKernel_Coa", suf, (nq, nq, nq), nelmt, b, inp)
        assert G.fe.last_backend() == "lanes"
    finally:
        G.fe.set_backend("auto")
    assert np.array_equal(oracle.from_coa(got, nelmt, nq ** 3), want_em)
    got = G.run_hex("BwdTransHexKernel_Coa", suf, (nq, nq, nq), nelmt, b, inp)
    assert G.fe.last_backend() == ("tpe" if suf == "f64" and nq == 5 else "coa-pipe" if suf == "f64" and nq in (8, 10) else "lanes")
    assert np.array_equal(oracle.from_coa(got, nelmt, nq ** 3), want_em)


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("backend,nqs", [("tpe", [2, 3, 4, 6, 8, 10]), ("rows", [4, 8, 11, 16, 20, 32])])
def test_quad_other_interleaved_backends_stay_bit_exact(G, suf, backend, nqs):
    for nq in nqs:
        nelmt = 32 * 3
        b0, b1, inp, want_em = quad_case(G, suf, nq, nelmt, 3200 + nq)
        try:
            G.fe.set_backend(backend)
            got = G.run_quad("BwdTransQuadKernel_Coa", suf, nq, nq, nelmt, b0, b1, inp)
            assert G.fe.last_backend() == ("rows-coa" if backend == "rows" else backend)
        finally:
            G.fe.set_backend("auto")
        assert np.array_equal(oracle.from_coa(got, nelmt, nq * nq), want_em), (backend, nq)


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("backend,nqs", [("tpe", [2, 3, 4, 5]), ("rows", [4, 6, 8, 11, 12])])
def test_hex_other_interleaved_backends_stay_bit_exact(G, suf, backend, nqs):
    for nq in nqs:
        nelmt = 32 * 2
        b, inp, want_em = hex_case(G, suf, nq, nelmt, 3300 + nq)
        try:
            G.fe.set_backend(backend)
            got = G.run_hex("BwdTransHexKernel_Coa", suf, (nq, nq, nq), nelmt, b, inp)
            assert G.fe.last_backend() == ("rows-coa" if backend == "rows" else backend)
        finally:
            G.fe.set_backend("auto")
        assert np.array_equal(oracle.from_coa(got, nelmt, nq ** 3), want_em), (backend, nq)


def test_lanes_backend_where_it_has_no_instantiation(G):
    try:
        G.fe.set_backend("lanes")
        with pytest.raises(Exception):  # interleaved: nq = 4 .. 16 and 32 only
            G.run_quad("BwdTransQuadKernel_Coa", "f64", 20, 20, 32, *quad_case(G, "f64", 20, 32, 2)[:3])
        b0, b1, inp, _ = quad_case(G, "f32", 9, 32, 1)
        with pytest.raises(Exception):  # element-major quads: even nq = 4 .. 16 only
            G.run_quad("BwdTransQuadKernel", "f32", 9, 9, 32, b0, b1, oracle.from_coa(inp, 32, 64))
        b, inph, _ = hex_case(G, "f64", 8, 32, 3)
        with pytest.raises(Exception):  # element-major hexes: nq = 4, 6 (and 8, 10 in FP32) only
            G.run_hex("BwdTransHexKernel", "f64", (8, 8, 8), 32, b, inph)
    finally:
        G.fe.set_backend("auto")


LANES_EM = [("f64", nq) for nq in (4, 6, 8, 10, 12, 14, 16)] + [("f32", nq) for nq in (4, 6, 8, 10, 12, 14, 16)]


@pytest.mark.parametrize("suf,nq", LANES_EM)
@pytest.mark.parametrize("nelmt", [1, 37, 160, 20011])
def test_quad_element_major_lanes_kernel_bit_exact(G, suf, nq, nelmt):
    """"lanes-em": the element-major even-nq quads that route to the bulk-copied-slab kernel; whole and ragged
    tiles (the ragged tail of the slab is fetched with plain loads), forced and through the default routing"""
    dt, nm = G.NP[suf], nq - 1
    rng = np.random.default_rng(3400 + nq + nelmt)
    b0, b1 = rnd(rng, nm * nq, dt), rnd(rng, nm * nq, dt)
    inp = rnd(rng, nelmt * nm * nm, dt)
    want = oracle.bwdtrans_quad(nq, nq, nelmt, b0, b1, inp)
    try:
        G.fe.set_backend("lanes")
        got = G.run_quad("BwdTransQuadKernel_QP_Shared", suf, nq, nq, nelmt, b0, b1, inp)
        assert G.fe.last_backend() == "lanes-em"
    finally:
        G.fe.set_backend("auto")
    assert np.array_equal(got, want)
    got = G.run_quad("BwdTransQuadKernel", suf, nq, nq, nelmt, b0, b1, inp)
    assert G.fe.last_backend() == "lanes-em"
    assert np.array_equal(got, want)


HEX_LANES_EM = [("f64", 4), ("f64", 6), ("f32", 4), ("f32", 6), ("f32", 8), ("f32", 10)]


@pytest.mark.parametrize("suf,nq", HEX_LANES_EM)
@pytest.mark.parametrize("nelmt", [1, 37, 160, 5003])
def test_hex_element_major_lanes_kernel_bit_exact(G, suf, nq, nelmt):
    dt, nm = G.NP[suf], nq - 1
    rng = np.random.default_rng(3600 + nq + nelmt)
    b = [rnd(rng, nm * nq, dt) for _ in range(3)]
    inp = rnd(rng, nelmt * nm ** 3, dt)
    want = oracle.bwdtrans_hex(nq, nq, nq, nelmt, *b, inp)
    try:
        G.fe.set_backend("lanes")
        got = G.run_hex("BwdTransHexKernel_QP_Shared", suf, (nq, nq, nq), nelmt, b, inp)
        assert G.fe.last_backend() == "lanes-em"
    finally:
        G.fe.set_backend("auto")
    assert np.array_equal(got, want)
    got = G.run_hex("BwdTransHexKernel", suf, (nq, nq, nq), nelmt, b, inp)
    assert G.fe.last_backend() == "lanes-em"
    assert np.array_equal(got, want)


@pytest.mark.parametrize("suf,nq", [("f64", 8), ("f32", 14)])
def test_quad_element_major_lanes_kernel_needs_an_aligned_slab(G, suf, nq):
    """the slab is fetched with a bulk copy (16-byte aligned source): a misaligned `in` takes the table's back-end
    under the default routing and is refused when lanes is forced; either way nothing is written outside `out`"""
    import torch
    dt, nm, nelmt = G.NP[suf], nq - 1, 333
    tdt = torch.float64 if suf == "f64" else torch.float32
    rng = np.random.default_rng(3500 + nq)
    b0, b1 = rnd(rng, nm * nq, dt), rnd(rng, nm * nq, dt)
    inp = rnd(rng, nelmt * nm * nm, dt)
    want = oracle.bwdtrans_quad(nq, nq, nelmt, b0, b1, inp)
    big_in = torch.zeros(nelmt * nm * nm + 4, dtype=tdt, device="cuda")
    big_in[1:1 + inp.size] = G.dev(inp)
    d_b0, d_b1 = G.dev(b0), G.dev(b1)
    d_out = torch.full((nelmt * nq * nq + 2,), float("nan"), dtype=tdt, device="cuda")
    isz = inp.itemsize
    st = torch.cuda.current_stream().cuda_stream
    G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", suf, nq, nq, nelmt, d_b0.data_ptr(), d_b1.data_ptr(),
                       big_in.data_ptr() + isz, d_out.data_ptr() + isz, stream=st)
    assert G.fe.last_backend() != "lanes-em"
    got = G.host(d_out)
    G.assert_parity(got[1:-1], want, suf)
    assert np.isnan(got[0]) and np.isnan(got[-1])
    try:
        G.fe.set_backend("lanes")
        with pytest.raises(Exception):
            G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", suf, nq, nq, nelmt, d_b0.data_ptr(), d_b1.data_ptr(),
                               big_in.data_ptr() + isz, d_out.data_ptr() + isz, stream=st)
    finally:
        G.fe.set_backend("auto")
    # aligned `in`, misaligned `out` (scalar stores: any alignment)
    d_in = G.dev(inp)
    d_out.fill_(float("nan"))
    G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", suf, nq, nq, nelmt, d_b0.data_ptr(), d_b1.data_ptr(),
                       d_in.data_ptr(), d_out.data_ptr() + isz, stream=st)
    assert G.fe.last_backend() == "lanes-em"
    got = G.host(d_out)
    assert np.array_equal(got[1:-1], want)
    assert np.isnan(got[0]) and np.isnan(got[-1])


@pytest.mark.parametrize("suf", ["f64", "f32"])
def test_lanes_full_size_checksum_of_checksums(G, suf):
    """BASELINE.json's size (64 Mi points) through a size-independent property: every element gets the same modes,
    so every element's output must equal element 0's, which the oracle checks; the layout makes lane e of group g
    a different address pattern for every (g, e)."""
    import torch
    dt = G.NP[suf]
    for dim, nq in ((2, 16), (3, 8)):
        nm = nq - 1
        nelmt = (64 << 20) // nq ** dim // 32 * 32
        rng = np.random.default_rng(77 + nq)
        b = [rnd(rng, nm * nq, dt) for _ in range(dim)]
        one = rnd(rng, nm ** dim, dt)
        d_in = G.dev(one).repeat_interleave(32).reshape(1, -1).repeat(nelmt // 32, 1).reshape(-1).contiguous()
        d_b = [G.dev(x) for x in b]
        d_out = torch.empty(nelmt * nq ** dim, dtype=d_in.dtype, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        if dim == 2:
            want = oracle.bwdtrans_quad(nq, nq, 1, b[0], b[1], one)
            G.fe.bwdtrans_quad("BwdTransQuadKernel_Coa", suf, nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                               d_in.data_ptr(), d_out.data_ptr(), stream=st)
        else:
            want = oracle.bwdtrans_hex(nq, nq, nq, 1, *b, one)
            G.fe.bwdtrans_hex("BwdTransHexKernel_Coa", suf, nq, nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                              d_b[2].data_ptr(), d_in.data_ptr(), d_out.data_ptr(), stream=st)
        assert G.fe.last_backend() == ("coa-pipe" if (dim, nq, suf) == (3, 8, "f64") else "lanes")
        torch.cuda.synchronize()
        # out_coa[g][m][e] == want[m] for every g, e
        view = d_out.reshape(nelmt // 32, nq ** dim, 32)
        d_want = G.dev(want).reshape(1, -1, 1)
        assert bool((view == d_want).all()), (dim, nq, suf)


@pytest.mark.parametrize("case", ["quad-em-f64-8", "quad-em-f32-14", "hex-em-f32-8", "hex-em-f64-6", "quad-coa-f64-12",
                                  "hex-coa-f32-6", "quad-em-f64-9", "hex-em-f64-10", "quad-em-f32-2", "hex-em-f64-2",
                                  "quad-coa-f32-4", "hex-coa-f64-3", "quad-coa-f64-20", "hex-coa-f64-10"])
def test_back_to_back_calls_with_alternating_bases(G, case):
    """Every call refills the per-device constant bank and the operator kernel that follows must see the new
    matrices.  (Regression: the operator used to be launched as a programmatic dependent of the fill; ptxas hoists the
    uniform loads of the basis above griddepcontrol.wait, so once in a few thousand calls a CTA multiplied by the
    previous call's basis.)  Small grids, 300 calls, the two bases alternate: a stale read cannot hide."""
    import torch
    kind, layout, suf, nq = case.split("-")
    nq = int(nq)
    dim = 2 if kind == "quad" else 3
    dt, nm = G.NP[suf], nq - 1
    nelmt = 64
    rng = np.random.default_rng(4000 + nq)
    bases = [[rnd(rng, nm * nq, dt) for _ in range(dim)] for _ in range(2)]
    inp_em = rnd(rng, nelmt * nm ** dim, dt)
    coa = layout == "coa"
    inp = oracle.to_coa(inp_em, nelmt, nm ** dim) if coa else inp_em
    want = []
    for bs in bases:
        w = (oracle.bwdtrans_quad(nq, nq, nelmt, *bs, inp_em) if dim == 2
             else oracle.bwdtrans_hex(nq, nq, nq, nelmt, *bs, inp_em))
        want.append(oracle.to_coa(w, nelmt, nq ** dim) if coa else w)
    d_b = [[G.dev(x) for x in bs] for bs in bases]
    d_in = G.dev(inp)
    reps = 300
    outs = [torch.empty(nelmt * nq ** dim, dtype=d_in.dtype, device="cuda") for _ in range(reps)]
    st = torch.cuda.current_stream().cuda_stream
    name = ("BwdTransQuadKernel" if dim == 2 else "BwdTransHexKernel") + ("_Coa" if coa else "_QP_Shared")
    # a second stream keeps every SM busy meanwhile, so the one-CTA fill kernel is not the only thing scheduled
    side = torch.cuda.Stream()
    noise = torch.ones(64 << 20, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for r in range(reps):
        if r % 10 == 0:
            with torch.cuda.stream(side):
                noise.mul_(1.0000001)
        b = d_b[r & 1]
        if dim == 2:
            G.fe.bwdtrans_quad(name, suf, nq, nq, nelmt, b[0].data_ptr(), b[1].data_ptr(), d_in.data_ptr(),
                               outs[r].data_ptr(), stream=st)
        else:
            G.fe.bwdtrans_hex(name, suf, nq, nq, nq, nelmt, b[0].data_ptr(), b[1].data_ptr(), b[2].data_ptr(),
                              d_in.data_ptr(), outs[r].data_ptr(), stream=st)
    torch.cuda.synchronize()
    d_want = [G.dev(w) for w in want]
    wrong = [r for r in range(reps) if not torch.equal(outs[r], d_want[r & 1])]
    assert not wrong, (case, G.fe.last_backend(), len(wrong), wrong[:10])


@pytest.mark.parametrize("dim,suf,nq", [(2, "f64", 16), (2, "f32", 14), (3, "f64", 6), (3, "f32", 8), (3, "f32", 10)])
def test_element_major_lanes_kernels_at_baseline_size(G, dim, suf, nq):
    """BASELINE.json's 64 Mi quadrature points through size-independent properties: identical elements give identical
    outputs (element 0 is checked against the oracle), the operator is exactly homogeneous under a power-of-two
    scale, and the fused checksum equals the checksum of what was stored."""
    import torch
    dt, nm = G.NP[suf], nq - 1
    nelmt = (64 << 20) // nq ** dim - 7            # not a multiple of any tile size
    rng = np.random.default_rng(5000 + nq)
    b = [rnd(rng, nm * nq, dt) for _ in range(dim)]
    one = rnd(rng, nm ** dim, dt)
    want = (oracle.bwdtrans_quad(nq, nq, 1, b[0], b[1], one) if dim == 2 else oracle.bwdtrans_hex(nq, nq, nq, 1, *b, one))
    d_b = [G.dev(x) for x in b]
    d_in = G.dev(one).repeat(nelmt)
    d_out = torch.empty(nelmt * nq ** dim, dtype=d_in.dtype, device="cuda")
    d_ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    scratch = torch.empty(G.fe.sumsq_scratch_bytes(), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    G.fe.bwdtrans_sumsq(suf, (nq,) * dim, nelmt, [x.data_ptr() for x in d_b], d_in.data_ptr(), d_out.data_ptr(),
                        d_ss.data_ptr(), scratch.data_ptr(), st)
    assert G.fe.last_backend() == "lanes-em"
    torch.cuda.synchronize()
    view = d_out.reshape(nelmt, nq ** dim)
    assert bool((view == G.dev(want).reshape(1, -1)).all())
    ss_want = float((view[0].double() ** 2).sum()) * nelmt
    assert abs(float(d_ss.item()) - ss_want) <= 1e-10 * ss_want
    # homogeneity: scaling the modes by 4 scales every output by exactly 4 (no rounding in a power-of-two scale)
    d_in.mul_(4)
    d_out4 = torch.empty_like(d_out)
    if dim == 2:
        G.fe.bwdtrans_quad("BwdTransQuadKernel", suf, nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                           d_in.data_ptr(), d_out4.data_ptr(), stream=st)
    else:
        G.fe.bwdtrans_hex("BwdTransHexKernel", suf, nq, nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                          d_b[2].data_ptr(), d_in.data_ptr(), d_out4.data_ptr(), stream=st)
    assert G.fe.last_backend() == "lanes-em"
    assert bool(torch.equal(d_out4, d_out * 4))


@pytest.mark.parametrize("layout,dim,suf,nq", [("em", 2, "f64", 12), ("em", 3, "f32", 8), ("coa", 2, "f32", 16),
                                               ("coa", 3, "f64", 7)])  # (coa FP64 nq = 8: test_coa_large_nq_gpu.py)
def test_lanes_kernels_keep_non_finite_values_inside_their_element(G, layout, dim, suf, nq):
    """one poisoned element (Inf modes) must not leak into its neighbours: every lane / tile slot is private"""
    dt, nm = G.NP[suf], nq - 1
    nelmt = 96
    rng = np.random.default_rng(5100 + nq)
    b = [rnd(rng, nm * nq, dt) for _ in range(dim)]
    inp = rnd(rng, nelmt * nm ** dim, dt).reshape(nelmt, -1)
    bad = 37
    inp[bad] = np.inf
    inp = inp.reshape(-1)
    want = (oracle.bwdtrans_quad(nq, nq, nelmt, b[0], b[1], inp) if dim == 2
            else oracle.bwdtrans_hex(nq, nq, nq, nelmt, *b, inp)).reshape(nelmt, -1)
    coa = layout == "coa"
    src = oracle.to_coa(inp, nelmt, nm ** dim) if coa else inp
    if dim == 2:
        got = G.run_quad("BwdTransQuadKernel_Coa" if coa else "BwdTransQuadKernel", suf, nq, nq, nelmt, b[0], b[1], src)
    else:
        got = G.run_hex("BwdTransHexKernel_Coa" if coa else "BwdTransHexKernel", suf, (nq,) * 3, nelmt, b, src)
    assert G.fe.last_backend() == ("lanes" if coa else "lanes-em")
    got = (oracle.from_coa(got, nelmt, nq ** dim) if coa else got).reshape(nelmt, -1)
    ok = [e for e in range(nelmt) if e != bad]
    assert np.isfinite(got[ok]).all()
    assert np.array_equal(got[ok], want[ok])
    assert not np.isfinite(got[bad]).any()

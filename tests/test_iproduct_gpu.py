"""IProductWRTBase through the C ABI.

The operator is not in the reference, so it has no golden vector: the oracle's restatement is pinned to the
(pinned) BwdTrans through the adjoint identity <IProduct(u), c> == <w*u, BwdTrans(c)>
(tests/test_oracle_golden.py::test_iproduct_oracle_is_the_adjoint_of_bwdtrans, CPU), and the CUDA kernels are
held bit for bit to that oracle here."""
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


def run(G, suf, dim, nq, nelmt, b, inp, w):
    import torch
    tdt = torch.float64 if suf == "f64" else torch.float32
    d_b = [G.dev(x) for x in b]
    d_in = G.dev(inp)
    d_w = G.dev(w) if w is not None else None
    nm = nq - 1
    d_out = torch.full((nelmt * nm ** dim + 2,), float("nan"), dtype=tdt, device="cuda")
    G.fe.iproduct(suf, (nq,) * dim, nelmt, [x.data_ptr() for x in d_b], d_in.data_ptr(),
                  d_out.data_ptr() + d_out.element_size(), weights=d_w.data_ptr() if d_w is not None else 0,
                  stream=torch.cuda.current_stream().cuda_stream)
    got = G.host(d_out)
    assert np.isnan(got[0]) and np.isnan(got[-1])          # nothing written outside
    return got[1:-1]


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("nq", [2, 3, 4, 5, 6, 8, 10, 12, 14, 16, 32])
def test_quad_bit_exact_with_oracle(G, suf, weighted, nq):
    dt = G.NP[suf]
    rng = np.random.default_rng(1100 + nq)
    nm = nq - 1
    b = [rng.standard_normal(nm * nq).astype(dt) for _ in range(2)]
    for nelmt in (1, 77, 2049):
        inp = rng.standard_normal(nelmt * nq * nq).astype(dt)
        w = rng.random(nelmt * nq * nq).astype(dt) + dt(0.5) if weighted else None
        want = oracle.iproduct_quad(nq, nq, nelmt, b[0], b[1], inp, w)
        got = run(G, suf, 2, nq, nelmt, b, inp, w)
        assert np.array_equal(got, want), (nq, nelmt, G.rel_max(got, want))


@pytest.mark.parametrize("suf", ["f64", "f32"])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("nq", [2, 3, 4, 5, 6, 8, 10, 12, 15])
def test_hex_bit_exact_with_oracle(G, suf, weighted, nq):
    dt = G.NP[suf]
    rng = np.random.default_rng(1200 + nq)
    nm = nq - 1
    b = [rng.standard_normal(nm * nq).astype(dt) for _ in range(3)]
    for nelmt in (1, 37, 300) if nq <= 10 else (1, 19):
        inp = rng.standard_normal(nelmt * nq ** 3).astype(dt)
        w = rng.random(nelmt * nq ** 3).astype(dt) + dt(0.5) if weighted else None
        want = oracle.iproduct_hex(nq, nq, nq, nelmt, *b, inp, w)
        got = run(G, suf, 3, nq, nelmt, b, inp, w)
        assert np.array_equal(got, want), (nq, nelmt, G.rel_max(got, want))


@pytest.mark.parametrize("backend", ["rows", "mma"])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("dim,nq", [(2, 8), (2, 12), (2, 14), (2, 16), (2, 32), (3, 8)])
def test_fp64_backends_bit_exact_ragged_groups_and_alignment(G, backend, weighted, dim, nq):
    """row kernel and tensor-core kernel forced in turn; group sizes that do not divide nelmt; 8-byte aligned slabs"""
    import torch
    nm = nq - 1
    rng = np.random.default_rng(1300 + nq)
    b = [rng.standard_normal(nm * nq) for _ in range(dim)]
    for nelmt, shift in ((1, 0), (13, 1), (1001, 0), (1001, 1)):
        if dim == 2 and nq == 32 and nelmt > 100:
            nelmt = 333
        inp = rng.standard_normal(nelmt * nq ** dim)
        w = rng.random(nelmt * nq ** dim) + 0.5 if weighted else None
        big_in = torch.zeros(inp.size + 2, dtype=torch.float64, device="cuda")
        big_in[shift:shift + inp.size] = torch.from_numpy(inp).cuda()
        big_w = None
        if weighted:
            big_w = torch.zeros(inp.size + 2, dtype=torch.float64, device="cuda")
            big_w[1 - shift:1 - shift + inp.size] = torch.from_numpy(w).cuda()   # the metric at the OTHER alignment
        d_b = [G.dev(x) for x in b]
        d_out = torch.full((nelmt * nm ** dim + 2,), float("nan"), dtype=torch.float64, device="cuda")
        try:
            G.fe.set_backend(backend)
            G.fe.iproduct("f64", (nq,) * dim, nelmt, [x.data_ptr() for x in d_b], big_in.data_ptr() + 8 * shift,
                          d_out.data_ptr() + 8, weights=(big_w.data_ptr() + 8 * (1 - shift)) if weighted else 0)
            assert G.fe.last_backend() == ("iprod-mma" if backend == "mma" else "iprod-rows")
        finally:
            G.fe.set_backend("auto")
        got = G.host(d_out)
        want = (oracle.iproduct_quad(nq, nq, nelmt, b[0], b[1], inp, w) if dim == 2
                else oracle.iproduct_hex(nq, nq, nq, nelmt, *b, inp, w))
        assert np.isnan(got[0]) and np.isnan(got[-1])
        assert np.array_equal(got[1:-1], want), (backend, nq, nelmt, shift, G.rel_max(got[1:-1], want))


@pytest.mark.parametrize("dim,nq,nelmt", [(2, 8, 1048576), (3, 8, 131072), (2, 12, 466033), (3, 6, 310689)])
def test_adjoint_of_the_device_bwdtrans_at_baseline_size(G, dim, nq, nelmt):
    """<IProduct(u), c> == <u, BwdTrans(c)> with both operators on the device, 64 Mi quadrature points"""
    import torch
    nm = nq - 1
    b = torch.randn(nm * nq, dtype=torch.float64, device="cuda")
    u = torch.randn(nelmt * nq ** dim, dtype=torch.float64, device="cuda")
    c = torch.randn(nelmt * nm ** dim, dtype=torch.float64, device="cuda")
    ip = torch.empty_like(c)
    bt = torch.empty_like(u)
    st = torch.cuda.current_stream().cuda_stream
    G.fe.iproduct("f64", (nq,) * dim, nelmt, [b.data_ptr()] * dim, u.data_ptr(), ip.data_ptr(), stream=st)
    if dim == 2:
        G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", "f64", nq, nq, nelmt, b.data_ptr(), b.data_ptr(), c.data_ptr(),
                           bt.data_ptr(), stream=st)
    else:
        G.fe.bwdtrans_hex("BwdTransHexKernel_QP_Shared", "f64", nq, nq, nq, nelmt, b.data_ptr(), b.data_ptr(),
                          b.data_ptr(), c.data_ptr(), bt.data_ptr(), stream=st)
    lhs, rhs = float(torch.dot(ip, c)), float(torch.dot(u, bt))
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), abs(rhs), float(torch.linalg.norm(ip) * torch.linalg.norm(c)) * 1e-2)


def test_argument_errors(G):
    import torch
    t = torch.zeros(4096, dtype=torch.float64, device="cuda")
    p = t.data_ptr()
    with pytest.raises(G.fe.B200feError) as e:   # unequal nq: outside the instantiated shapes
        G.fe.lib().b200fe_IProductWRTBaseQuad_f64  # symbol exists
        G.fe.iproduct("f64", (4, 6), 8, [p, p], p, p)
    assert e.value.code == G.fe.E_UNSUPPORTED
    with pytest.raises(G.fe.B200feError) as e:
        G.fe.iproduct("f64", (4, 4), 8, [p, 0], p, p)
    assert e.value.code == G.fe.E_INVAL
    G.fe.iproduct("f64", (4, 4, 4), 0, [p, p, p], p, p)  # empty: no-op


IPL_QUAD = [(suf, nq) for suf in ("f64", "f32") for nq in (4, 6, 8, 10, 12, 14, 16)]
IPL_HEX = [("f64", 4), ("f64", 6), ("f64", 8), ("f32", 4), ("f32", 6), ("f32", 8), ("f32", 10)]


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("dim,suf,nq", [(2, s, n) for s, n in IPL_QUAD] + [(3, s, n) for s, n in IPL_HEX])
def test_lanes_kernel_forced_bit_exact(G, dim, suf, nq, weighted):
    """sumfac_iprod_lanes.cuh, every instantiation, whole and ragged tiles; a thread fetches its own row / plane with
    8- or 16-byte loads, so a slab that is not 16-byte aligned is refused when forced and takes the row kernel under
    the default routing"""
    import torch
    dt, nm = G.NP[suf], nq - 1
    tdt = torch.float64 if suf == "f64" else torch.float32
    rng = np.random.default_rng(1500 + nq + dim)
    b = [rng.standard_normal(nm * nq).astype(dt) for _ in range(dim)]
    for nelmt in (1, 37, 1001):
        inp = rng.standard_normal(nelmt * nq ** dim).astype(dt)
        w = (rng.random(nelmt * nq ** dim) + 0.5).astype(dt) if weighted else None
        want = (oracle.iproduct_quad(nq, nq, nelmt, b[0], b[1], inp, w) if dim == 2
                else oracle.iproduct_hex(nq, nq, nq, nelmt, *b, inp, w))
        try:
            G.fe.set_backend("lanes")
            got = run(G, suf, dim, nq, nelmt, b, inp, w)
            assert G.fe.last_backend() == "iprod-lanes"
        finally:
            G.fe.set_backend("auto")
        assert np.array_equal(got, want), (dim, suf, nq, nelmt, G.rel_max(got, want))
    # misaligned input slab
    nelmt = 37
    inp = rng.standard_normal(nelmt * nq ** dim).astype(dt)
    want = (oracle.iproduct_quad(nq, nq, nelmt, b[0], b[1], inp, None) if dim == 2
            else oracle.iproduct_hex(nq, nq, nq, nelmt, *b, inp, None))
    big_in = torch.zeros(inp.size + 4, dtype=tdt, device="cuda")
    big_in[1:1 + inp.size] = G.dev(inp)
    d_b = [G.dev(x) for x in b]
    d_out = torch.full((nelmt * nm ** dim,), float("nan"), dtype=tdt, device="cuda")
    G.fe.iproduct(suf, (nq,) * dim, nelmt, [x.data_ptr() for x in d_b], big_in.data_ptr() + inp.itemsize, d_out.data_ptr())
    assert G.fe.last_backend() != "iprod-lanes"
    assert np.array_equal(G.host(d_out), want)
    try:
        G.fe.set_backend("lanes")
        with pytest.raises(G.fe.B200feError):
            G.fe.iproduct(suf, (nq,) * dim, nelmt, [x.data_ptr() for x in d_b], big_in.data_ptr() + inp.itemsize,
                          d_out.data_ptr())
    finally:
        G.fe.set_backend("auto")


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("suf,nq", [("f64", 8), ("f64", 10), ("f32", 8), ("f32", 10)])
def test_hex_pipe_kernel_bit_exact_and_default_route(G, suf, nq, weighted):
    """sumfac_iprod.cuh, persistent TMA-fed twin of the hex row kernel (round 2): forced and through the default
    routing; one tile, ragged last tile, more tiles than resident CTAs (every CTA loops and refills its ring); a slab
    that is not 16-byte aligned leaves the route (bulk copies need it) and is refused when forced; nothing is written
    outside `out`"""
    import torch
    dt, nm = G.NP[suf], nq - 1
    tdt = torch.float64 if suf == "f64" else torch.float32
    rng = np.random.default_rng(1700 + nq)
    b = [rng.standard_normal(nm * nq).astype(dt) for _ in range(3)]
    d_b = [G.dev(x) for x in b]
    for nelmt in (1, 5, 1001, 148 * 8 * 4 * 2 + 3):
        inp = rng.standard_normal(nelmt * nq ** 3).astype(dt)
        w = (rng.random(nelmt * nq ** 3) + 0.5).astype(dt) if weighted else None
        want = oracle.iproduct_hex(nq, nq, nq, nelmt, *b, inp, w)
        d_in, d_w = G.dev(inp), (G.dev(w) if weighted else None)
        for forced in (True, False):
            d_out = torch.full((nelmt * nm ** 3 + 2,), float("nan"), dtype=tdt, device="cuda")
            try:
                if forced:
                    G.fe.set_backend("pipe")
                G.fe.iproduct(suf, (nq,) * 3, nelmt, [x.data_ptr() for x in d_b], d_in.data_ptr(),
                              d_out.data_ptr() + inp.itemsize, weights=d_w.data_ptr() if weighted else 0)
                default_is_pipe = not (suf == "f64" and nq == 8 and weighted)  # that case measured faster on the row kernel
                assert G.fe.last_backend() == ("iprod-pipe" if forced or default_is_pipe else "iprod-rows")
            finally:
                G.fe.set_backend("auto")
            got = G.host(d_out)
            assert np.isnan(got[0]) and np.isnan(got[-1])
            assert np.array_equal(got[1:-1], want), (suf, nq, nelmt, forced, G.rel_max(got[1:-1], want))
    # misaligned input slab
    nelmt = 37
    inp = rng.standard_normal(nelmt * nq ** 3).astype(dt)
    want = oracle.iproduct_hex(nq, nq, nq, nelmt, *b, inp, None)
    big_in = torch.zeros(inp.size + 4, dtype=tdt, device="cuda")
    big_in[1:1 + inp.size] = G.dev(inp)
    d_out = torch.full((nelmt * nm ** 3,), float("nan"), dtype=tdt, device="cuda")
    G.fe.iproduct(suf, (nq,) * 3, nelmt, [x.data_ptr() for x in d_b], big_in.data_ptr() + inp.itemsize, d_out.data_ptr())
    assert G.fe.last_backend() == "iprod-rows"
    assert np.array_equal(G.host(d_out), want)
    try:
        G.fe.set_backend("pipe")
        with pytest.raises(G.fe.B200feError):
            G.fe.iproduct(suf, (nq,) * 3, nelmt, [x.data_ptr() for x in d_b], big_in.data_ptr() + inp.itemsize, d_out.data_ptr())
        with pytest.raises(G.fe.B200feError):  # no instantiation at other nq, none for quads
            G.fe.iproduct(suf, (6,) * 3, 8, [x.data_ptr() for x in d_b], d_out.data_ptr(), d_out.data_ptr())
    finally:
        G.fe.set_backend("auto")

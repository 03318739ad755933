"""b200fe_plan_*: the basis matrices uploaded once for many operator calls (include/b200fe.h, an extension over the
reference's per-launch basis arguments, benchmark04.cc:912-1001 / benchmark05.cc:1262-1385).

A plan call must be bit-identical to the per-call entry point (same kernels, same bank contents), must skip the
staging launch only while the bank really holds its matrices -- another plan, a per-call entry point or an
IProductWRTBase (transposed bank) in between forces a refill -- and must keep working after the caller frees its
own copy of the matrices.
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


def oracle_bwd(dim, nq, nelmt, bs, inp_em):
    return (oracle.bwdtrans_quad(nq, nq, nelmt, *bs, inp_em) if dim == 2
            else oracle.bwdtrans_hex(nq, nq, nq, nelmt, *bs, inp_em))


# (dim, suf, nq): bank-fed back-ends (lanes-em, lanes, tpe, rows, pipe) and the tensor-core ones (basis from global)
CASES = [(2, "f64", 4), (2, "f64", 7), (2, "f64", 16), (2, "f64", 32), (2, "f32", 2), (2, "f32", 8), (2, "f32", 13),
         (2, "f32", 32), (3, "f64", 2), (3, "f64", 6), (3, "f64", 8), (3, "f64", 10), (3, "f32", 5), (3, "f32", 8)]


@pytest.mark.parametrize("dim,suf,nq", CASES)
@pytest.mark.parametrize("coa", [False, True], ids=["em", "coa"])
def test_plan_matches_the_oracle_and_the_per_call_entry_point(G, dim, suf, nq, coa):
    import torch
    dt, nm = G.NP[suf], nq - 1
    nelmt = 96
    rng = np.random.default_rng(8000 + 100 * dim + nq)
    bs = [rnd(rng, nm * nq, dt) for _ in range(dim)]
    inp_em = rnd(rng, nelmt * nm ** dim, dt)
    want = oracle_bwd(dim, nq, nelmt, bs, inp_em)
    inp = oracle.to_coa(inp_em, nelmt, nm ** dim) if coa else inp_em
    want = oracle.to_coa(want, nelmt, nq ** dim) if coa else want
    d_b = [G.dev(x) for x in bs]
    d_in = G.dev(inp)
    st = torch.cuda.current_stream().cuda_stream
    plan = G.fe.Plan(dim, suf, nq, [b.data_ptr() for b in d_b], stream=st)
    try:
        # the plan owns a copy: the caller's matrices may go away (poison them first, then drop them)
        for b in d_b:
            b.fill_(float("nan"))
        torch.cuda.synchronize()
        outs = []
        for _ in range(3):
            d_out = torch.full((nelmt * nq ** dim,), float("nan"), dtype=d_in.dtype, device="cuda")
            plan.bwdtrans(nelmt, d_in.data_ptr(), d_out.data_ptr(), coa=coa, stream=st)
            outs.append(G.host(d_out))
        G.assert_parity(outs[0], want, suf, ("plan", dim, suf, nq, coa))
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
        # the per-call entry point on fresh copies of the matrices gives the same bits
        if dim == 2:
            ref = G.run_quad("BwdTransQuadKernel_Coa" if coa else "BwdTransQuadKernel_QP_Shared", suf, nq, nq, nelmt,
                             bs[0], bs[1], inp)
        else:
            ref = G.run_hex("BwdTransHexKernel_Coa" if coa else "BwdTransHexKernel_QP_Shared", suf, (nq, nq, nq),
                            nelmt, bs, inp)
        assert np.array_equal(outs[0], ref)
    finally:
        plan.destroy()


@pytest.mark.parametrize("dim,suf,nq", [(2, "f64", 8), (2, "f32", 6), (3, "f64", 6), (3, "f32", 8)])
def test_staging_launch_is_skipped_only_while_the_bank_is_this_plans(G, dim, suf, nq):
    """Launch counts (b200fe_launch_count): the first call of a plan stages (2 launches), repeats do not (1), and
    anything else that used the bank in between -- another plan, a per-call entry point, the plan's own
    IProductWRTBase (transposed bank) -- costs a restage.  Every output is checked, so a skipped fill that should not
    have been skipped shows as a wrong result."""
    import torch
    dt, nm = G.NP[suf], nq - 1
    nelmt = 64
    rng = np.random.default_rng(8500 + 100 * dim + nq)
    bsets = [[rnd(rng, nm * nq, dt) for _ in range(dim)] for _ in range(3)]
    inp = rnd(rng, nelmt * nm ** dim, dt)
    wants = [oracle_bwd(dim, nq, nelmt, bs, inp) for bs in bsets]
    d_bs = [[G.dev(x) for x in bs] for bs in bsets]
    d_in = G.dev(inp)
    d_want = [G.dev(w) for w in wants]
    st = torch.cuda.current_stream().cuda_stream
    plans = [G.fe.Plan(dim, suf, nq, [b.data_ptr() for b in d_bs[i]], stream=st) for i in range(2)]

    def call_plan(i):
        d_out = torch.full((nelmt * nq ** dim,), float("nan"), dtype=d_in.dtype, device="cuda")
        before = G.fe.launch_count()
        plans[i].bwdtrans(nelmt, d_in.data_ptr(), d_out.data_ptr(), stream=st)
        launches = G.fe.launch_count() - before
        torch.cuda.synchronize()
        assert torch.equal(d_out, d_want[i]), ("plan", i, G.fe.last_backend())
        return launches

    def call_plain():
        d_out = torch.full((nelmt * nq ** dim,), float("nan"), dtype=d_in.dtype, device="cuda")
        b = d_bs[2]
        if dim == 2:
            G.fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", suf, nq, nq, nelmt, b[0].data_ptr(), b[1].data_ptr(),
                               d_in.data_ptr(), d_out.data_ptr(), stream=st)
        else:
            G.fe.bwdtrans_hex("BwdTransHexKernel_QP_Shared", suf, nq, nq, nq, nelmt, b[0].data_ptr(), b[1].data_ptr(),
                              b[2].data_ptr(), d_in.data_ptr(), d_out.data_ptr(), stream=st)
        torch.cuda.synchronize()
        assert torch.equal(d_out, d_want[2]), "per-call entry point"

    try:
        assert call_plan(0) == 2           # stage + operator
        assert call_plan(0) == 1           # resident
        assert call_plan(0) == 1
        assert call_plan(1) == 2           # the other plan takes the bank
        assert call_plan(0) == 2           # ... and loses it again
        call_plain()                       # anonymous fill
        assert call_plan(0) == 2
        assert call_plan(0) == 1
        # the plan's IProductWRTBase uses the transposed bank layout: BwdTrans after it must restage
        q_in = G.dev(rnd(rng, nelmt * nq ** dim, dt))
        m_out = torch.empty(nelmt * nm ** dim, dtype=d_in.dtype, device="cuda")
        plans[0].iproduct(nelmt, q_in.data_ptr(), m_out.data_ptr(), stream=st)
        ref_out = torch.empty_like(m_out)
        G.fe.iproduct(suf, (nq,) * dim, nelmt, [b.data_ptr() for b in d_bs[0]], q_in.data_ptr(), ref_out.data_ptr(),
                      stream=st)
        torch.cuda.synchronize()
        assert torch.equal(m_out, ref_out)
        assert call_plan(0) == 2
        # a second stream of the same device: ordered behind the fill by the bank's event, still no restage
        side = torch.cuda.Stream()
        d_out = torch.full((nelmt * nq ** dim,), float("nan"), dtype=d_in.dtype, device="cuda")
        torch.cuda.synchronize()
        before = G.fe.launch_count()
        plans[0].bwdtrans(nelmt, d_in.data_ptr(), d_out.data_ptr(), stream=side.cuda_stream)
        assert G.fe.launch_count() - before == 1
        side.synchronize()
        assert torch.equal(d_out, d_want[0])
    finally:
        for p in plans:
            p.destroy()


def test_plan_argument_errors(G):
    import torch
    fe = G.fe
    b = G.dev(np.ones(7 * 8))
    x = G.dev(np.ones(64 * 49))
    out = torch.empty(64 * 64, dtype=torch.float64, device="cuda")
    with pytest.raises(fe.B200feError) as e:
        fe.Plan(4, "f64", 8, [b.data_ptr(), b.data_ptr(), b.data_ptr()])
    assert e.value.code == fe.E_INVAL
    with pytest.raises(fe.B200feError) as e:
        fe.Plan(2, "f64", 33, [b.data_ptr(), b.data_ptr()])
    assert e.value.code == fe.E_UNSUPPORTED
    with pytest.raises(fe.B200feError) as e:
        fe.Plan(3, "f64", 8, [b.data_ptr(), b.data_ptr()])          # basis2 missing
    assert e.value.code == fe.E_INVAL
    plan = fe.Plan(2, "f64", 8, [b.data_ptr(), b.data_ptr()])
    try:
        with pytest.raises(fe.B200feError) as e:
            plan.bwdtrans(48, x.data_ptr(), out.data_ptr(), coa=True)   # interleaved layout needs whole groups
        assert e.value.code == fe.E_INVAL
        with pytest.raises(fe.B200feError) as e:
            plan.bwdtrans(64, 0, out.data_ptr())
        assert e.value.code == fe.E_INVAL
        with pytest.raises(fe.B200feError) as e:
            plan.bwdtrans(64, x.data_ptr() + 4, out.data_ptr())
        assert e.value.code == fe.E_ALIGN
        plan.bwdtrans(0, x.data_ptr(), out.data_ptr())                  # empty range: nothing to do
    finally:
        plan.destroy()
    plan.destroy()                                                      # idempotent on the wrapper


@pytest.mark.parametrize("suf", ["f64", "f32"])
def test_documented_bank_fill_route_is_bit_identical(G, suf):
    """b200fe_set_bank_fill("memcpy"): the basis reaches the constant bank through cudaMemcpyToSymbolAsync instead of
    the fill kernel's stores + programmatic dependent launch (include/b200fe.h).  Same kernels, same bits -- with two
    different bases alternating, so a stale bank would show."""
    rng = np.random.default_rng(77)
    dt = G.NP[suf]
    cases = [("quad", 4, 4096), ("quad", 10, 2048), ("hex", 6, 1024)]
    try:
        for mode in ("memcpy", "kernel"):
            G.fe.set_bank_fill(mode)
            for op, nq, nelmt in cases:
                nm, dim = nq - 1, 2 if op == "quad" else 3
                for rep in range(3):
                    b = [rng.standard_normal(nm * nq).astype(dt) for _ in range(dim)]
                    inp = rng.standard_normal(nelmt * nm ** dim).astype(dt)
                    if op == "quad":
                        got = G.run_quad("BwdTransQuadKernel_QP_Shared", suf, nq, nq, nelmt, b[0], b[1], inp)
                        want = oracle.bwdtrans_quad(nq, nq, nelmt, b[0], b[1], inp)
                    else:
                        got = G.run_hex("BwdTransHexKernel_QP_Shared", suf, (nq,) * 3, nelmt, b, inp)
                        want = oracle.bwdtrans_hex(nq, nq, nq, nelmt, b[0], b[1], b[2], inp)
                    assert np.array_equal(got, want), (mode, op, nq, rep, G.fe.last_backend())
    finally:
        G.fe.set_bank_fill("kernel")

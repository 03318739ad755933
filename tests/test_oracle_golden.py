"""Pin the CPU oracle to the reference's own known-answer data.

The reference's golden vectors are the `norm:` columns of its committed logs
(tests/golden/ref_norms.json, made by tests/golden/make_golden.py).  They are
printed with setprecision(10), so they pin results to ~5e-10 relative.
"""
import math

import numpy as np
import pytest

import oracle

PRINT_TOL = 6e-10  # 10 significant digits


def rel(a, b):
    return abs(a - b) / abs(b)


QUAD_NQ = [2, 4, 6, 8, 10, 12, 14, 16, 32]
HEX_NQ = [2, 4, 6, 8, 10]


@pytest.mark.parametrize("nq", QUAD_NQ)
@pytest.mark.parametrize("coa", [False, True])
def test_quad_norms_match_reference_logs(golden, nq, coa):
    nm = nq - 1
    b = oracle.gen_basis(nm, nq)
    for nelmt in (128, 1024):
        inp = oracle.gen_in(nelmt, nm * nm, coa=coa)
        out = oracle.bwdtrans_quad(nq, nq, nelmt, b, b, inp, coa=coa, use_fma=True)
        got = math.sqrt(oracle.sumsq(out))
        ref = golden["quad"][str(nq)][str(nelmt)]
        for col, want in enumerate(ref):
            assert rel(got, want) < PRINT_TOL, (nq, nelmt, col, got, want)


@pytest.mark.parametrize("nq", HEX_NQ)
@pytest.mark.parametrize("coa", [False, True])
def test_hex_norms_match_reference_logs(golden, nq, coa):
    nm = nq - 1
    b = oracle.gen_basis(nm, nq)
    nelmt = 128
    inp = oracle.gen_in(nelmt, nm ** 3, coa=coa)
    out = oracle.bwdtrans_hex(nq, nq, nq, nelmt, b, b, b, inp, coa=coa, use_fma=True)
    got = math.sqrt(oracle.sumsq(out))
    ref = golden["hex"][str(nq)][str(nelmt)]
    bad = set(golden["hex_invalid_columns"])
    for col, want in enumerate(ref):
        if col in bad:
            continue  # reference bug benchmark05.cc:193, see SURVEY.md 2.2
        assert rel(got, want) < PRINT_TOL, (nq, col, got, want)


def test_hex_full_size_norm_by_scaling(golden):
    # all elements are identical, so norm(nelmt) = norm(128) * sqrt(nelmt/128):
    # checks the 1 Mi-element golden of the headline case without 4 GB of data
    nq, nm = 8, 7
    b = oracle.gen_basis(nm, nq)
    out = oracle.bwdtrans_hex(nq, nq, nq, 128, b, b, b, oracle.gen_in(128, nm ** 3))
    got = math.sqrt(oracle.sumsq(out) * (1048576 / 128))
    assert rel(got, golden["hex"]["8"]["1048576"][0]) < PRINT_TOL


def test_fma_and_plain_rounding_agree_to_tolerance():
    # nvcc contracts `tmp += a*b` to FMA, g++ does not: both roundings must sit
    # inside the 1e-12 parity tolerance north_star states for FP64
    nq, nm, nelmt = 8, 7, 64
    b = oracle.gen_basis(nm, nq)
    inp = oracle.gen_in(nelmt, nm ** 3)
    a = oracle.bwdtrans_hex(nq, nq, nq, nelmt, b, b, b, inp, use_fma=True)
    c = oracle.bwdtrans_hex(nq, nq, nq, nelmt, b, b, b, inp, use_fma=False)
    scale = np.abs(a).max()
    assert np.abs(a - c).max() / scale < 1e-12


@pytest.mark.parametrize("nq", [3, 4, 7])
def test_layouts_are_consistent(nq):
    # interleaved result == re-laid-out element-major result, bit for bit, on
    # element-dependent data (identical elements would hide index bugs)
    nm, nelmt = nq - 1, 96
    rng = np.random.default_rng(7)
    b0, b1, b2 = (rng.standard_normal(nm * nq) for _ in range(3))
    inp = rng.standard_normal(nelmt * nm ** 3)
    ref = oracle.bwdtrans_hex(nq, nq, nq, nelmt, b0, b1, b2, inp)
    coa = oracle.bwdtrans_hex(nq, nq, nq, nelmt, b0, b1, b2, oracle.to_coa(inp, nelmt, nm ** 3), coa=True)
    assert np.array_equal(oracle.from_coa(coa, nelmt, nq ** 3), ref)
    inq = rng.standard_normal(nelmt * nm ** 2)
    refq = oracle.bwdtrans_quad(nq, nq, nelmt, b0, b1, inq)
    coaq = oracle.bwdtrans_quad(nq, nq, nelmt, b0, b1, oracle.to_coa(inq, nelmt, nm ** 2), coa=True)
    assert np.array_equal(oracle.from_coa(coaq, nelmt, nq ** 2), refq)


def test_quad_matches_dense_tensor_contraction():
    # independent restatement: out[e,j,i] = sum_qp in[e,q,p] B0[p,i] B1[q,j]
    nq0, nq1, nelmt = 5, 7, 10
    nm0, nm1 = nq0 - 1, nq1 - 1
    rng = np.random.default_rng(3)
    b0, b1 = rng.standard_normal(nm0 * nq0), rng.standard_normal(nm1 * nq1)
    inp = rng.standard_normal(nelmt * nm0 * nm1)
    got = oracle.bwdtrans_quad(nq0, nq1, nelmt, b0, b1, inp).reshape(nelmt, nq1, nq0)
    want = np.einsum("eqp,pi,qj->eji", inp.reshape(nelmt, nm1, nm0), b0.reshape(nm0, nq0), b1.reshape(nm1, nq1))
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_hex_matches_dense_tensor_contraction():
    nq0, nq1, nq2, nelmt = 3, 4, 5, 6
    nm0, nm1, nm2 = nq0 - 1, nq1 - 1, nq2 - 1
    rng = np.random.default_rng(5)
    b0, b1, b2 = (rng.standard_normal(a * b) for a, b in ((nm0, nq0), (nm1, nq1), (nm2, nq2)))
    inp = rng.standard_normal(nelmt * nm0 * nm1 * nm2)
    got = oracle.bwdtrans_hex(nq0, nq1, nq2, nelmt, b0, b1, b2, inp).reshape(nelmt, nq2, nq1, nq0)
    want = np.einsum("erqp,pi,qj,rk->ekji", inp.reshape(nelmt, nm2, nm1, nm0), b0.reshape(nm0, nq0),
                     b1.reshape(nm1, nq1), b2.reshape(nm2, nq2))
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


# ---- benchmark01-03 ---------------------------------------------------------------

B_SIZES = [1024, 2048, 65536, 1048576, 4194304]


@pytest.mark.parametrize("n", B_SIZES)
def test_b01_norm(golden, n):
    for fused in (True, False):  # device rounding of benchmark01.cc:178 (FMA-contracted by nvcc) / host rounding
        x = oracle.set_data(n, fused=fused)
        got = math.sqrt(oracle.sumsq(x))
        for want in golden["b01"][str(n)]:
            assert rel(got, want) < PRINT_TOL


def test_b01_fused_generator_differs_from_host_rounding_by_at_most_one_ulp():
    n = 300000
    a, b = oracle.set_data(n, fused=True), oracle.set_data(n, fused=False)
    assert not np.array_equal(a, b)            # the contraction is observable ...
    assert np.abs(a - b).max() <= np.spacing(np.abs(b).max())  # ... in the last bit only


def test_b01_generator_is_bit_exact_integer_modulo():
    n = 300000
    i = np.arange(n, dtype=np.uint32)
    want = (i % 13).astype(np.float64) + (0.2 + 0.00001 * (i % 100191).astype(np.float64))
    assert np.array_equal(oracle.set_data(n), want)
    want2 = (i % 8).astype(np.float64) + (0.4 + 0.00003 * (i % 100721).astype(np.float64))
    assert np.array_equal(oracle.set_data(n, second=True), want2)


@pytest.mark.parametrize("n", B_SIZES)
def test_b02_norm_after_40_in_place_adds(golden, n):
    x = oracle.set_data(n)
    y = oracle.set_data(n, second=True)
    oracle.add_vector(x, y, reps=40)
    got = math.sqrt(oracle.sumsq(x))
    for want in golden["b02"][str(n)]:
        assert rel(got, want) < PRINT_TOL


def test_b02_add_is_elementwise_exact():
    n = 1000
    x, y = oracle.set_data(n), oracle.set_data(n, second=True)
    want = x.copy()
    for _ in range(3):
        want = want + y
    assert np.array_equal(oracle.add_vector(x, y, reps=3), want)


@pytest.mark.parametrize("n", [128, 256, 1024, 2048])
def test_b03_norm(golden, n):
    A, x = oracle.gen_matvec(n, n)
    y = oracle.matvec(n, n, A, x)
    got = math.sqrt(oracle.sumsq(y))
    for want in golden["b03"][str(n)]:
        # different GPU libraries sum rows in different orders; the log's own
        # columns agree to the printed digits
        assert rel(got, want) < 5e-9


def test_f32_variants_run_and_track_f64():
    nq, nm, nelmt = 4, 3, 64
    b64, b32 = oracle.gen_basis(nm, nq), oracle.gen_basis(nm, nq, np.float32)
    i64, i32 = oracle.gen_in(nelmt, nm * nm), oracle.gen_in(nelmt, nm * nm, np.float32)
    o64 = oracle.bwdtrans_quad(nq, nq, nelmt, b64, b64, i64)
    o32 = oracle.bwdtrans_quad(nq, nq, nelmt, b32, b32, i32)
    assert o32.dtype == np.float32
    assert np.abs(o32 - o64).max() / np.abs(o64).max() < 1e-5


# ---- IProductWRTBase: not in the reference, pinned through the adjoint identity --------------------

@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-13), (np.float32, 2e-6)])
@pytest.mark.parametrize("nq", [2, 3, 4, 7, 8, 10])
def test_iproduct_oracle_is_the_adjoint_of_bwdtrans(dt, tol, nq):
    """<IProduct(u), c> == <w*u, BwdTrans(c)> for random u, c, w: ties the unpinned operator to the pinned one"""
    rng = np.random.default_rng(5000 + nq)
    nm, nelmt = nq - 1, 9
    b = [rng.standard_normal(nm * nq).astype(dt) for _ in range(3)]
    for dim in (2, 3):
        u = rng.standard_normal(nelmt * nq ** dim).astype(dt)
        c = rng.standard_normal(nelmt * nm ** dim).astype(dt)
        w = (rng.random(nelmt * nq ** dim) + 0.5).astype(dt)
        if dim == 2:
            ip = oracle.iproduct_quad(nq, nq, nelmt, b[0], b[1], u, w)
            bt = oracle.bwdtrans_quad(nq, nq, nelmt, b[0], b[1], c)
        else:
            ip = oracle.iproduct_hex(nq, nq, nq, nelmt, *b, u, w)
            bt = oracle.bwdtrans_hex(nq, nq, nq, nelmt, *b, c)
        lhs = np.dot(ip.astype(np.float64), c.astype(np.float64))
        rhs = np.dot((u * w).astype(np.float64), bt.astype(np.float64))
        scale = np.linalg.norm(ip.astype(np.float64)) * np.linalg.norm(c.astype(np.float64))
        assert abs(lhs - rhs) <= tol * scale


def test_iproduct_oracle_known_answer():
    """nq = 2 by hand: one mode per direction, out = sum_ij B0[i] B1[j] w in"""
    b0, b1 = np.array([2.0, 3.0]), np.array([5.0, 7.0])
    x = np.array([1.0, 10.0, 100.0, 1000.0])     # [j][i]
    w = np.array([1.0, 0.5, 0.25, 2.0])
    want = sum(b0[i] * b1[j] * x[2 * j + i] * w[2 * j + i] for i in range(2) for j in range(2))
    got = oracle.iproduct_quad(2, 2, 1, b0, b1, x, w)
    assert got.shape == (1,) and got[0] == want

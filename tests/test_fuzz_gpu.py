"""Randomised differential test: every entry point x dtype x nq x element count x slab alignment against the
oracle, default routing (so whatever back-end the tables pick is what gets exercised)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

QUAD = ["BwdTransQuadKernel", "BwdTransQuadKernel_Coa", "BwdTransQuadKernel_QP", "BwdTransQuadKernel_QP_Shared",
        "BwdTransQuadKernel_QP_1D", "BwdTransQuadKernel_QP_1D_Shared"]
HEX = [k.replace("Quad", "Hex") for k in QUAD]


@pytest.fixture(scope="module")
def G():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests import gpu_util
    return gpu_util


@pytest.mark.parametrize("seed", range(12))
def test_random_cases(G, seed):
    import torch
    rng = np.random.default_rng(90210 + seed)
    seen = set()
    for _ in range(25):
        dim = int(rng.integers(2, 4))
        suf = ["f64", "f32"][int(rng.integers(0, 2))]
        dt = G.NP[suf]
        tdt = torch.float64 if suf == "f64" else torch.float32
        nq = int(rng.integers(2, 33 if dim == 2 else 13))
        nm = nq - 1
        kernel = (QUAD if dim == 2 else HEX)[int(rng.integers(0, 6))]
        coa = kernel.endswith("_Coa")
        cap = 4000 if nq <= 10 else (600 if dim == 2 else 80)
        nelmt = int(rng.integers(1, cap))
        if coa:
            nelmt = max(32, nelmt // 32 * 32)
        shift = 0 if coa else int(rng.integers(0, 4))      # element offset of the slabs: any sizeof(T) alignment
        b = [rng.standard_normal(nm * nq).astype(dt) for _ in range(dim)]
        inp = rng.standard_normal(nelmt * nm ** dim).astype(dt)
        src = oracle.to_coa(inp, nelmt, nm ** dim) if coa else inp
        big_in = torch.zeros(src.size + 4, dtype=tdt, device="cuda")
        big_in[shift:shift + src.size] = torch.from_numpy(src).cuda()
        nout = nelmt * nq ** dim
        big_out = torch.full((nout + 8,), float("nan"), dtype=tdt, device="cuda")
        d_b = [G.dev(x) for x in b]
        isz = big_in.element_size()
        wsp = torch.empty(max(1, nelmt * nq ** (dim - 1) * nm * nq), dtype=tdt, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        if dim == 2:
            G.fe.bwdtrans_quad(kernel, suf, nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(),
                               big_in.data_ptr() + isz * shift, big_out.data_ptr() + isz * (shift + 1),
                               wsp=wsp.data_ptr(), stream=st)
            want = oracle.bwdtrans_quad(nq, nq, nelmt, b[0], b[1], inp)
        else:
            G.fe.bwdtrans_hex(kernel, suf, nq, nq, nq, nelmt, d_b[0].data_ptr(), d_b[1].data_ptr(), d_b[2].data_ptr(),
                              big_in.data_ptr() + isz * shift, big_out.data_ptr() + isz * (shift + 1),
                              wsp0=wsp.data_ptr(), wsp1=wsp.data_ptr(), stream=st)
            want = oracle.bwdtrans_hex(nq, nq, nq, nelmt, *b, inp)
        seen.add(G.fe.last_backend())
        got = G.host(big_out)
        body = got[shift + 1: shift + 1 + nout]
        if coa:
            body = oracle.from_coa(body, nelmt, nq ** dim)
        G.assert_parity(body, want, suf, (kernel, suf, nq, nelmt, shift))
        assert np.isnan(got[:shift + 1]).all() and np.isnan(got[shift + 1 + nout:]).all(), "wrote outside out"
    assert seen  # at least something ran

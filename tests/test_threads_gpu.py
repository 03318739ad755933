"""The C ABI under several host threads ("safe to call from N host threads with N different
devices/streams" -- what utils/multi_gpu.h does with one thread per GPU).

The bank-fed back-ends share one constant bank per device and dtype: fill -> operator -> release is a critical
section per device, and a fill on another stream waits for the previous user's event.  Two threads that alternate
DIFFERENT basis matrices through the same bank on two streams must therefore each get their own result, bit for bit
with the oracle, on every call; threads on different devices hold different locks and different banks.
ctypes drops the GIL for the duration of a foreign call, so the calls below really overlap.
"""
import threading

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


def make_case(G, kind, layout, suf, nq, nelmt, seed):
    """host input in the entry point's layout, the two basis sets a thread alternates, and the oracle's answers"""
    dim = 2 if kind == "quad" else 3
    dt, nm = G.NP[suf], nq - 1
    rng = np.random.default_rng(seed)
    bases = [[rnd(rng, nm * nq, dt) for _ in range(dim)] for _ in range(2)]
    inp_em = rnd(rng, nelmt * nm ** dim, dt)
    coa = layout == "coa"
    want = []
    for bs in bases:
        w = (oracle.bwdtrans_quad(nq, nq, nelmt, *bs, inp_em) if dim == 2
             else oracle.bwdtrans_hex(nq, nq, nq, nelmt, *bs, inp_em))
        want.append(oracle.to_coa(w, nelmt, nq ** dim) if coa else w)
    inp = oracle.to_coa(inp_em, nelmt, nm ** dim) if coa else inp_em
    name = ("BwdTransQuadKernel" if dim == 2 else "BwdTransHexKernel") + ("_Coa" if coa else "_QP_Shared")
    return dict(dim=dim, suf=suf, nq=nq, nelmt=nelmt, bases=bases, inp=inp, want=want, name=name)


def worker(G, case, device, reps, start, errors, slot):
    """`reps` calls on an own stream of `device`, bases alternating; every output is kept and compared afterwards"""
    import torch
    try:
        torch.cuda.set_device(device)
        dev = torch.device("cuda", device)
        put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        d_b = [[put(x) for x in bs] for bs in case["bases"]]
        d_in = put(case["inp"])
        d_want = [put(w) for w in case["want"]]
        nq, dim, nelmt = case["nq"], case["dim"], case["nelmt"]
        outs = [torch.empty(nelmt * nq ** dim, dtype=d_in.dtype, device=dev) for _ in range(reps)]
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.synchronize(dev)
        start.wait()
        for r in range(reps):
            b = d_b[r & 1]
            if dim == 2:
                G.fe.bwdtrans_quad(case["name"], case["suf"], nq, nq, nelmt, b[0].data_ptr(), b[1].data_ptr(),
                                   d_in.data_ptr(), outs[r].data_ptr(), stream=stream.cuda_stream)
            else:
                G.fe.bwdtrans_hex(case["name"], case["suf"], nq, nq, nq, nelmt, b[0].data_ptr(), b[1].data_ptr(),
                                  b[2].data_ptr(), d_in.data_ptr(), outs[r].data_ptr(), stream=stream.cuda_stream)
        stream.synchronize()
        wrong = [r for r in range(reps) if not torch.equal(outs[r], d_want[r & 1])]
        if wrong:
            errors[slot] = ("wrong results", case["name"], case["suf"], nq, len(wrong), wrong[:10])
    except Exception as exc:  # a thread's exception must fail the test, not vanish
        errors[slot] = ("exception", repr(exc))


def run_threads(G, cases, devices, reps):
    start = threading.Barrier(len(cases))
    errors = [None] * len(cases)
    threads = [threading.Thread(target=worker, args=(G, c, d, reps, start, errors, i))
               for i, (c, d) in enumerate(zip(cases, devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a worker thread hangs"
    assert errors == [None] * len(cases), errors


# every pair shares ONE bank (same dimension and dtype => same translation unit of the library)
SAME_BANK_PAIRS = [
    ("quad-em-f64-8", "quad-coa-f64-6"),     # lanes-em and lanes
    ("quad-em-f32-14", "quad-coa-f32-4"),    # lanes-em and thread-per-element
    ("hex-em-f64-6", "hex-coa-f64-4"),       # lanes-em and lanes
    ("hex-em-f32-8", "hex-em-f32-5"),        # lanes-em and rows / pipe
]


@pytest.mark.parametrize("pair", SAME_BANK_PAIRS, ids=lambda p: "+".join(p))
def test_two_threads_two_streams_share_a_bank(G, pair):
    cases = []
    for i, spec in enumerate(pair):
        kind, layout, suf, nq = spec.split("-")
        cases.append(make_case(G, kind, layout, suf, int(nq), 64, 7000 + 10 * i + int(nq)))
    run_threads(G, cases, [0, 0], reps=200)


def test_one_thread_per_device(G):
    """utils/multi_gpu.h's pattern: each host thread drives its own device through the same entry point"""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("one visible device")
    ndev = min(ndev, 8)
    cases = [make_case(G, "hex", "em", "f64", 6, 64, 7100 + d) for d in range(ndev)]
    run_threads(G, cases, list(range(ndev)), reps=200)

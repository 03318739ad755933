"""The non-headline legs of bench.py: per-operator sweeps (configs[2]/[3]), the benchmark01-03 sweep (configs[1]) and
the same-box baselines (the reference's own kernels compiled for sm_100a in oracle/_ref, and the reference's cuBLAS
formulation in oracle/libref_blas.so) timed on the same B200 next to every row.

Everything here runs OUTSIDE the headline's timed region.  The two oracle-side libraries are measurement baselines
only: they are timed, never used to produce a libb200fe result.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so")
BLAS_SO = os.path.join(ROOT, "oracle", "libref_blas.so")
REF_VARIANTS = ["Uncoales", "Coales", "QP", "QP/Shared", "QP-1D", "QP-1D/Shared"]  # columns 6-11 of the reference logs

QUAD_NQ = (2, 4, 6, 8, 10, 12, 14, 16, 32)   # benchmark04/run.sh:5
HEX_NQ = (2, 4, 6, 8, 10)                    # benchmark05/run.sh:5
POINTS = 1 << 26                             # "~64 M quadrature points" (BASELINE.json configs[2]/[3])


def gen_basis(nm, nq, dtype="float64"):
    """B[k] = cos((T)k), k = p*nq + i -- the reference's synthetic basis (benchmark05.cc:1216-1236)"""
    return np.cos(np.arange(nm * nq, dtype=dtype)).astype(dtype)


def gen_in(nelmt, nmtot, dtype="float64"):
    """in[e][k] = sin((T)(k+1)) for every element (benchmark05.cc:1195-1215), element-major"""
    one = np.sin(np.arange(1, nmtot + 1, dtype=dtype)).astype(dtype)
    return np.tile(one, nelmt)


def alg_bytes_per_elem(dim, nq, size):
    return size * ((nq - 1) ** dim + nq ** dim)


def nelmt_for(dim, nq):
    return max(32, (POINTS // nq ** dim) // 32 * 32)


def time_min(torch, fn, reps=5, warm=3):
    """min over reps of the CUDA-event time around one call (current stream)"""
    for _ in range(warm):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * reps)]
    for r in range(reps):
        ev[2 * r].record()
        fn()
        ev[2 * r + 1].record()
    torch.cuda.synchronize()
    return min(ev[2 * r].elapsed_time(ev[2 * r + 1]) for r in range(reps))


def time_mean(torch, fn, reps=10, warm=1):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _vp(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else None)


class SameBox:
    """the reference's kernels (oracle/_ref, launch shapes of run_test: benchmark04.cc:907-1012,
    benchmark05.cc:1260-1374, threads = 128, elblocks = 1) and its cuBLAS formulation, callable on torch buffers"""

    def __init__(self):
        self.ref = ctypes.CDLL(REF_SO) if os.path.exists(REF_SO) else None
        self.blas = ctypes.CDLL(BLAS_SO) if os.path.exists(BLAS_SO) else None

    def available(self):
        return {"ref_kernels": self.ref is not None, "cublas": self.blas is not None}

    def quad(self, torch, suf, nq, nelmt, b, d_in, d_out, reps=3):
        nm = nq - 1
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        out = {}
        if self.ref is not None:
            w0 = torch.empty(nelmt * nm, dtype=d_in.dtype, device="cuda")
            w1 = torch.empty(nelmt * nq * nm, dtype=d_in.dtype, device="cuda")
            fn = getattr(self.ref, f"ref_bwdtrans_quad_{suf}")
            ms = []
            for v in range(6):
                def call(v=v):
                    rc = fn(ctypes.c_int(v), ctypes.c_uint(nq), ctypes.c_uint(nq), ctypes.c_uint(nelmt), _vp(b), _vp(b),
                            _vp(d_in), _vp(w0), _vp(w1), _vp(d_out), ctypes.c_uint(128), ctypes.c_uint(1), st)
                    assert rc == 0, (v, rc)
                ms.append(time_min(torch, call, reps=reps, warm=1))
            out["ref_ms"] = ms
            del w0, w1
        if self.blas is not None:
            w = torch.empty(nelmt * nq * nm, dtype=d_in.dtype, device="cuda")
            fn = getattr(self.blas, f"ref_cublas_bwdtrans_quad_{suf}")

            def call():
                rc = fn(ctypes.c_int(nq), ctypes.c_int(nq), ctypes.c_int(nelmt), _vp(b), _vp(b), _vp(d_in), _vp(w),
                        _vp(d_out), st)
                assert rc == 0, rc
            out["cublas_ms"] = time_min(torch, call, reps=reps, warm=2)
            del w
        return out

    def hex(self, torch, suf, nq, nelmt, b, d_in, d_out, reps=3):
        nm = nq - 1
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        out = {}
        if self.ref is not None:
            w0 = torch.empty(nelmt * nm * nm, dtype=d_in.dtype, device="cuda")
            w1 = torch.empty(nelmt * nm, dtype=d_in.dtype, device="cuda")
            w2 = torch.empty(nelmt * nq * nm * nm, dtype=d_in.dtype, device="cuda")
            w3 = torch.empty(nelmt * nq * nq * nm, dtype=d_in.dtype, device="cuda")
            fn = getattr(self.ref, f"ref_bwdtrans_hex_{suf}")
            ms = []
            for v in range(6):
                def call(v=v):
                    rc = fn(ctypes.c_int(v), ctypes.c_uint(nq), ctypes.c_uint(nq), ctypes.c_uint(nq), ctypes.c_uint(nelmt),
                            _vp(b), _vp(b), _vp(b), _vp(d_in), _vp(w0), _vp(w1), _vp(w2), _vp(w3), _vp(d_out),
                            ctypes.c_uint(128), ctypes.c_uint(1), st)
                    assert rc == 0, (v, rc)
                ms.append(time_min(torch, call, reps=reps, warm=1))
            out["ref_ms"] = ms
            del w0, w1, w2, w3
        if self.blas is not None:
            w1 = torch.empty(nelmt * nq * nm * nm, dtype=d_in.dtype, device="cuda")
            w2 = torch.empty(nelmt * nq * nq * nm, dtype=d_in.dtype, device="cuda")
            fn = getattr(self.blas, f"ref_cublas_bwdtrans_hex_{suf}")

            def call():
                rc = fn(ctypes.c_int(nq), ctypes.c_int(nq), ctypes.c_int(nq), ctypes.c_int(nelmt), _vp(b), _vp(b), _vp(b),
                        _vp(d_in), _vp(w1), _vp(w2), _vp(d_out), st)
                assert rc == 0, rc
            out["cublas_ms"] = time_min(torch, call, reps=reps, warm=2)
            del w1, w2
        return out


def sweep_operators(fe, torch, peak, same_box=None, reps=5, log=None):
    """configs[2]/[3]: every operator / nq / dtype at ~64 Mi quadrature points through the element-major entry point
    (`_QP_Shared`) and the interleaved one (`_Coa`); per-call and plan figures; the same-box baselines beside them"""
    rows = []
    st = torch.cuda.current_stream().cuda_stream
    for dim, nqs in ((2, QUAD_NQ), (3, HEX_NQ)):
        op = "quad" if dim == 2 else "hex"
        kern = "BwdTransQuadKernel" if dim == 2 else "BwdTransHexKernel"
        for suf, tdt, npdt, size in (("f64", torch.float64, np.float64, 8), ("f32", torch.float32, np.float32, 4)):
            for nq in nqs:
                nm = nq - 1
                nelmt = nelmt_for(dim, nq)
                b = torch.from_numpy(gen_basis(nm, nq, npdt)).cuda()
                one = torch.from_numpy(gen_in(32, nm ** dim, npdt)).cuda()
                d_in = one.view(32, -1).repeat(nelmt // 32, 1).reshape(-1).contiguous()
                d_out = torch.empty(nelmt * nq ** dim, dtype=tdt, device="cuda")
                byts = nelmt * alg_bytes_per_elem(dim, nq, size)
                modes = nelmt * nm ** dim

                def call(k):
                    if dim == 2:
                        fe.bwdtrans_quad(k, suf, nq, nq, nelmt, b.data_ptr(), b.data_ptr(), d_in.data_ptr(),
                                         d_out.data_ptr(), stream=st)
                    else:
                        fe.bwdtrans_hex(k, suf, nq, nq, nq, nelmt, b.data_ptr(), b.data_ptr(), b.data_ptr(),
                                        d_in.data_ptr(), d_out.data_ptr(), stream=st)

                row = {"op": op, "nq": nq, "dtype": suf, "nelmt": nelmt}
                for layout, k in (("em", kern + "_QP_Shared"), ("coa", kern + "_Coa")):
                    ms = time_min(torch, lambda: call(k), reps=reps)
                    row[layout] = {"backend": fe.last_backend(), "ms": round(ms, 4),
                                   "gdof_s": round(1e-6 * modes / ms, 2), "hbm_frac": round(1e-6 * byts / ms / peak, 4)}
                    try:  # the same call through a plan (b200fe_plan_*: basis staged once)
                        plan = fe.Plan(dim, suf, nq, [b.data_ptr()] * dim, stream=st)
                        pms = time_min(torch, lambda: plan.bwdtrans(nelmt, d_in.data_ptr(), d_out.data_ptr(),
                                                                    coa=(layout == "coa"), stream=st), reps=reps)
                        plan.destroy()
                        row[layout]["plan_ms"] = round(pms, 4)
                        row[layout]["plan_hbm_frac"] = round(1e-6 * byts / pms / peak, 4)
                    except Exception as exc:  # an auxiliary figure must not cost the row
                        row[layout]["plan_error"] = repr(exc)
                if same_box is not None:
                    try:
                        sb = (same_box.quad if dim == 2 else same_box.hex)(torch, suf, nq, nelmt, b, d_in, d_out)
                        if "ref_ms" in sb:
                            g = [round(1e-6 * modes / m, 2) for m in sb["ref_ms"]]
                            row["ref_gdof_s"] = g                                   # the six reference variants
                            em = [g[i] for i in (0, 2, 3, 4, 5)]
                            row["ref_best_em"] = max(em)
                            row["ref_best_em_variant"] = REF_VARIANTS[(0, 2, 3, 4, 5)[em.index(max(em))]]
                            row["ref_coa"] = g[1]
                        if "cublas_ms" in sb:
                            row["cublas_gdof_s"] = round(1e-6 * modes / sb["cublas_ms"], 2)
                    except Exception as exc:
                        row["same_box_error"] = repr(exc)
                rows.append(row)
                if log:
                    log("sweep", row)
                del d_in, d_out
    return rows


def sweep_iproduct(fe, torch, peak, reps=5):
    """IProductWRTBase at ~64 Mi quadrature points, without and with the quadrature metric w (one value
    per point: its nq^d values per element are algorithmic bytes of the weighted operator)"""
    out = []
    st = torch.cuda.current_stream().cuda_stream
    for dim, nqs in ((2, (4, 6, 8, 10, 12, 14, 16, 32)), (3, (4, 6, 8, 10))):
        for suf, tdt, size in (("f64", torch.float64, 8), ("f32", torch.float32, 4)):
            for nq in nqs:
                nm = nq - 1
                nelmt = nelmt_for(dim, nq)
                b = torch.from_numpy(gen_basis(nm, nq, "float64")).to(tdt).cuda()
                d_in = torch.randn(nelmt * nq ** dim, dtype=tdt, device="cuda")
                d_w = torch.rand(nelmt * nq ** dim, dtype=tdt, device="cuda") + 0.5
                d_out = torch.empty(nelmt * nm ** dim, dtype=tdt, device="cuda")
                rec = {"op": "iproduct_" + ("quad" if dim == 2 else "hex"), "nq": nq, "dtype": suf, "nelmt": nelmt}
                for weighted in (False, True):
                    try:
                        ms = time_min(torch, lambda weighted=weighted: fe.iproduct(
                            suf, (nq,) * dim, nelmt, [b.data_ptr()] * dim, d_in.data_ptr(), d_out.data_ptr(),
                            weights=d_w.data_ptr() if weighted else 0, stream=st), reps=reps)
                    except Exception as exc:  # a shape without an instantiation must not cost the line
                        rec["weighted" if weighted else "plain"] = {"error": repr(exc), "hbm_frac": 0.0}
                        continue
                    byts = nelmt * (alg_bytes_per_elem(dim, nq, size) + (size * nq ** dim if weighted else 0))
                    rec["weighted" if weighted else "plain"] = {
                        "backend": fe.last_backend(), "ms": round(ms, 4),
                        "gdof_s": round(1e-6 * nelmt * nm ** dim / ms, 2), "hbm_frac": round(1e-6 * byts / ms / peak, 4)}
                out.append(rec)
                del d_in, d_out, d_w
    return out


def sweep_fused(fe, torch, reps=10):
    """operator + sum(out^2) at ~64 Mi quadrature points: the two-pass form the reference uses after
    every variant against the fused entry point, for one shape per kernel family that carries the epilogue"""
    out = []
    st = torch.cuda.current_stream().cuda_stream
    d_scr = torch.empty(fe.sumsq_scratch_bytes(), dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(2, dtype=torch.float64, device="cuda")
    for dim, nq, suf, tdt in ((2, 8, "f32", torch.float32), (2, 16, "f64", torch.float64), (3, 6, "f64", torch.float64),
                              (3, 8, "f32", torch.float32), (3, 8, "f64", torch.float64)):
        nm = nq - 1
        nelmt = nelmt_for(dim, nq)
        b = torch.from_numpy(gen_basis(nm, nq, "float64")).to(tdt).cuda()
        d_in = torch.randn(nelmt * nm ** dim, dtype=tdt, device="cuda")
        d_out = torch.empty(nelmt * nq ** dim, dtype=tdt, device="cuda")

        def plain():
            if dim == 2:
                fe.bwdtrans_quad("BwdTransQuadKernel_QP_Shared", suf, nq, nq, nelmt, b.data_ptr(), b.data_ptr(),
                                 d_in.data_ptr(), d_out.data_ptr(), stream=st)
            else:
                fe.bwdtrans_hex("BwdTransHexKernel_QP_Shared", suf, nq, nq, nq, nelmt, b.data_ptr(), b.data_ptr(),
                                b.data_ptr(), d_in.data_ptr(), d_out.data_ptr(), stream=st)
            fe.sumsq(suf, d_out.data_ptr(), d_out.numel(), d_res.data_ptr(), d_scr.data_ptr(), st)

        def fused():
            fe.bwdtrans_sumsq(suf, (nq,) * dim, nelmt, [b.data_ptr()] * dim, d_in.data_ptr(), d_out.data_ptr(),
                              d_res.data_ptr() + 8, d_scr.data_ptr(), st)

        ms = [time_mean(torch, fn, reps=reps) for fn in (plain, fused)]
        r = d_res.cpu().numpy()
        out.append({"op": "quad" if dim == 2 else "hex", "nq": nq, "dtype": suf, "nelmt": nelmt,
                    "backend": fe.last_backend(), "operator_then_checksum_ms": round(ms[0], 4),
                    "fused_ms": round(ms[1], 4), "agree": bool(abs(r[0] - r[1]) <= 1e-12 * abs(r[0]))})
        del d_in, d_out
    return out


# ---- configs[1]: benchmark01-03 ------------------------------------------------------------------------------------

def sweep_vec(fe, torch, peak, same_box=None, max_log2=30, reps=5, log=None):
    """benchmark01 (sum x^2, 8 B/value), benchmark02 (x += y, 24 B/value) and benchmark03 (y = A x, 8*M*N B) over
    2^20 ... 2^30 doubles, GB/s with the reference's byte formulas (benchmark01.cc:330, benchmark02.cc:255,
    benchmark03.cc:332).  `dev_ms`: CUDA events around the kernels alone (the roofline figure); `ref_region_ms`
    (b01): host clock around the reference's whole timed region -- memset, memset, l2norm, reduce, D2H copy
    (benchmark01.cc:243-253) -- through the C ABI.  Sizes below 2^26 doubles are L2-resident or launch-bound and are
    reported without counting towards `min_frac`."""
    import time
    st = torch.cuda.current_stream().cuda_stream
    cst = ctypes.c_void_p(st)
    ref = same_box.ref if same_box is not None else None
    out = {"b01": [], "b02": [], "b03": []}
    f64 = torch.float64
    for lg in range(20, max_log2 + 1):
        n = 1 << lg
        blocks = min((n + 255) // 256, 1024)                       # the reference's grid (benchmark01.cc:236-238)
        x = torch.empty(n, dtype=f64, device="cuda")
        fe.set_data("f64", x.data_ptr(), n, stream=st)
        sums = torch.zeros(1024, dtype=f64, device="cuda")
        res = torch.zeros(1, dtype=f64, device="cuda")
        h_res = torch.zeros(1, dtype=f64).pin_memory()

        def l2(vl):
            fe.l2norm_vl("f64", sums.data_ptr(), x.data_ptr(), n, blocks, vl, stream=st)
            fe.reduce_vl("f64", res.data_ptr(), sums.data_ptr(), blocks, vl, stream=st)

        row = {"n": n, "log2": lg}
        for vl in (0, 1):
            ms = time_min(torch, lambda: l2(vl), reps=reps)
            row["vl" if vl else "scalar"] = {"dev_ms": round(ms, 5), "gb_s": round(8e-6 * n / ms, 1),
                                             "frac": round(8e-6 * n / ms / peak, 4)}
        # the reference's timed region, host clock (min of 10)
        best = None
        for _ in range(10):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sums.zero_()
            res.zero_()
            l2(1)
            h_res.copy_(res)                                       # synchronous D2H, as cudaMemcpy in the reference
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        row["ref_region_ms"] = round(1e3 * best, 5)
        row["ref_region_gb_s"] = round(8e-9 * n / best, 1)
        if ref is not None:
            rres = torch.zeros(1, dtype=f64, device="cuda")
            rs = torch.zeros(1024, dtype=f64, device="cuda")
            ms = time_min(torch, lambda: ref.ref_l2norm_f64(_vp(rres), _vp(rs), _vp(x), ctypes.c_uint(n),
                                                            ctypes.c_int(1), cst), reps=reps)
            row["ref_kernel_gb_s"] = round(8e-6 * n / ms, 1)
        out["b01"].append(row)
        if log:
            log("b01", row)

        y = torch.empty(n, dtype=f64, device="cuda")
        fe.set_data("f64", y.data_ptr(), n, second=True, stream=st)
        row = {"n": n, "log2": lg}
        for vl in (0, 1):
            ms = time_min(torch, lambda: fe.add_vector("f64", x.data_ptr(), y.data_ptr(), n, vl, stream=st), reps=reps)
            row["vl" if vl else "scalar"] = {"dev_ms": round(ms, 5), "gb_s": round(24e-6 * n / ms, 1),
                                             "frac": round(24e-6 * n / ms / peak, 4)}
        if ref is not None:
            ms = time_min(torch, lambda: ref.ref_add_vector_f64(_vp(x), _vp(y), ctypes.c_uint(n), ctypes.c_int(1), cst),
                          reps=reps)
            row["ref_kernel_gb_s"] = round(24e-6 * n / ms, 1)
        out["b02"].append(row)
        if log:
            log("b02", row)
        del x, y

        if lg % 2 == 0:                                            # square matrices with 2^lg values: size = 2^(lg/2)
            size = 1 << (lg // 2)
            A = torch.empty(size * size, dtype=f64, device="cuda")
            fe.set_data("f64", A.data_ptr(), size * size, stream=st)   # values do not matter for the timing
            xv = torch.arange(size, dtype=f64, device="cuda")
            yv = torch.empty(size, dtype=f64, device="cuda")
            row = {"size": size, "n": size * size, "log2": lg}
            for vl in (0, 1):
                ms = time_min(torch, lambda: fe.compute_matvec("f64", size, size, A.data_ptr(), xv.data_ptr(),
                                                               yv.data_ptr(), vl, stream=st), reps=reps)
                row["vl" if vl else "scalar"] = {"dev_ms": round(ms, 5), "gb_s": round(8e-6 * size * size / ms, 1),
                                                 "frac": round(8e-6 * size * size / ms / peak, 4)}
            if ref is not None:
                ms = time_min(torch, lambda: ref.ref_matvec_f64(ctypes.c_uint(size), ctypes.c_uint(size), _vp(A), _vp(xv),
                                                                _vp(yv), ctypes.c_int(1), cst), reps=reps)
                row["ref_kernel_gb_s"] = round(8e-6 * size * size / ms, 1)
            out["b03"].append(row)
            if log:
                log("b03", row)
            del A
    return out


# ---- compact summaries for the JSON line ---------------------------------------------------------------------------

def summarize_operators(rows, target=0.75):
    """what the driver's record has to show without profiles/: worst row, rows below the target, per-family fractions,
    same-box speed-ups.  nq < 4 is outside north_star's target and is listed separately."""
    def key(r, layout):
        return f"{r['op']}{'_coa' if layout == 'coa' else ''}/{r['dtype']}/nq{r['nq']}"
    fr = {}
    for r in rows:
        for layout in ("em", "coa"):
            fr[key(r, layout)] = max(r[layout]["hbm_frac"], r[layout].get("plan_hbm_frac", 0.0))
    in_target = {k: v for k, v in fr.items() if int(k.rsplit("nq", 1)[1]) >= 4}
    worst = min(in_target, key=in_target.get)
    fam = {}
    for r in rows:
        for layout in ("em", "coa"):
            fam.setdefault(f"{r['op']}{'_coa' if layout == 'coa' else ''}_{r['dtype']}", {})[str(r["nq"])] = \
                round(fr[key(r, layout)], 3)
    s = {"n_rows": len(fr), "n_rows_nq_ge_4": len(in_target), "points_per_row": POINTS, "target": target,
         "min_frac": round(in_target[worst], 4), "worst_row": worst,
         "rows_below_target": sorted(([k, round(v, 3)] for k, v in in_target.items() if v < target), key=lambda kv: kv[1]),
         "rows_nq_lt_4": {k: round(v, 3) for k, v in fr.items() if k not in in_target},
         "frac_by_family": fam, "basis": "best of the per-call entry point and the same call through a b200fe_plan"}
    sb_rows = [r for r in rows if "ref_best_em" in r or "cublas_gdof_s" in r]
    if sb_rows:
        sp_ref, sp_coa, sp_blas, table = {}, {}, {}, {}
        for r in sb_rows:
            ours_em = max(r["em"]["gdof_s"], 1e-6 * r["nelmt"] * (r["nq"] - 1) ** (2 if r["op"] == "quad" else 3)
                          / r["em"].get("plan_ms", 1e30))
            ours_coa = r["coa"]["gdof_s"]
            k = f"{r['op']}/{r['dtype']}/nq{r['nq']}"
            ent = {"ours": round(ours_em, 1), "ours_coa": round(ours_coa, 1)}
            if "ref_best_em" in r:
                sp_ref[k] = ours_em / r["ref_best_em"]
                sp_coa[k] = ours_coa / r["ref_coa"]
                ent.update({"ref_kernel": r["ref_best_em"], "ref_variant": r["ref_best_em_variant"], "ref_coa": r["ref_coa"],
                            "x_ref": round(sp_ref[k], 2), "x_ref_coa": round(sp_coa[k], 2)})
            if "cublas_gdof_s" in r:
                sp_blas[k] = ours_em / r["cublas_gdof_s"]
                ent.update({"cublas": r["cublas_gdof_s"], "x_cublas": round(sp_blas[k], 2)})
            table[k] = ent
        sb = {"unit": "GDoF/s", "what": "the reference's own kernels (oracle/_ref, its launch shapes, best of its five "
              "element-major variants / its Coales variant) and its cuBLAS formulation (benchmark04.cc:804-820, "
              "benchmark05.cc:1128-1153) timed on this GPU at the same sizes", "rows": table}
        for name, d in (("speedup_vs_ref_kernel", sp_ref), ("speedup_vs_ref_coa", sp_coa), ("speedup_vs_cublas", sp_blas)):
            if d:
                lo = min(d, key=d.get)
                sb[name] = {"min": round(d[lo], 2), "min_row": lo, "max": round(max(d.values()), 2),
                            "geomean": round(float(np.exp(np.mean(np.log(list(d.values()))))), 2)}
        s["same_box"] = sb
    return s


def summarize_iproduct(rows, target=0.75):
    fr = {f"{r['op']}/{r['dtype']}/nq{r['nq']}": max(r["plain"]["hbm_frac"], r["weighted"]["hbm_frac"]) for r in rows}
    worst = min(fr, key=fr.get)
    return {"n_rows": len(fr), "min_frac": round(fr[worst], 4), "worst_row": worst,
            "rows_below_target": sorted(([k, round(v, 3)] for k, v in fr.items() if v < target), key=lambda kv: kv[1]),
            "plain": {f"{r['op']}/{r['dtype']}/nq{r['nq']}": r["plain"]["hbm_frac"] for r in rows},
            "weighted": {f"{r['op']}/{r['dtype']}/nq{r['nq']}": r["weighted"]["hbm_frac"] for r in rows},
            "note": "parity-unpinned operator (absent from the reference); best of unweighted / weighted per row"}


def summarize_vec(v):
    s = {"unit": "GB/s", "sizes_log2": [r["log2"] for r in v["b01"]],
         "note": "device-timed kernels; frac counted for >= 2^26 doubles (smaller sizes are L2-resident or launch-bound)"}
    for b in ("b01", "b02", "b03"):
        rows = v[b]
        best = [max(r["vl"]["gb_s"], r["scalar"]["gb_s"]) for r in rows]
        fr = [max(r["vl"]["frac"], r["scalar"]["frac"]) for r in rows]
        big = [f for r, f in zip(rows, fr) if r["log2"] >= 26]
        e = {"gb_s": best, "frac_min_large": round(min(big), 4) if big else None,
             "frac_max_large": round(max(big), 4) if big else None}
        if b == "b03":
            e["sizes"] = [r["size"] for r in rows]
        if "ref_kernel_gb_s" in rows[0]:
            e["ref_kernel_gb_s"] = [r["ref_kernel_gb_s"] for r in rows]
            e["speedup_vs_ref_kernel_large"] = round(min(bb / r["ref_kernel_gb_s"] for r, bb in zip(rows, best)
                                                         if r["log2"] >= 26), 2) if big else None
        if b == "b01":
            e["ref_region_gb_s"] = [r["ref_region_gb_s"] for r in rows]
        s[b] = e
    return s

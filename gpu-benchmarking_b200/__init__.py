"""ctypes binding of libb200fe.so -- the C ABI declared in include/b200fe.h.

This is plumbing for tests/ and bench.py (the product is the shared library
and the C++ benchmark drivers).  Pointers are passed as integers
(`tensor.data_ptr()`), streams as `torch.cuda.current_stream().cuda_stream`.
There is no fallback: if the library is missing, import fails.
"""
import ctypes
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("B200FE_LIB") or os.path.join(_HERE, "libb200fe.so")  # (B200FE_LIB: an experimental build)
HEADER = os.path.join(ROOT, "include", "b200fe.h")

E_OK, E_INVAL, E_UNSUPPORTED, E_ALIGN, E_NODEVICE = 0, -1, -2, -3, -4


class B200feError(RuntimeError):
    def __init__(self, fn, code):
        self.code = code
        names = {E_INVAL: "B200FE_EINVAL", E_UNSUPPORTED: "B200FE_EUNSUPPORTED", E_ALIGN: "B200FE_EALIGN",
                 E_NODEVICE: "B200FE_ENODEVICE"}
        what = names.get(code, f"cudaError {code}" if code > 0 else str(code))
        super().__init__(f"{fn} failed: {what}")


def build(jobs=8):
    """compile libb200fe.so in-tree for sm_100a (nvcc cross-compiles without a GPU)"""
    subprocess.check_call(["make", "-s", "-j", str(jobs), "-C", os.path.join(_HERE, "csrc")])
    return LIB_PATH


def declared_symbols():
    """every entry point include/b200fe.h declares"""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200fe_\w+)\s*\(", text)))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU or PyTorch fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.b200fe_version.restype = ctypes.c_char_p
        _lib.b200fe_last_backend.restype = ctypes.c_char_p
        _lib.b200fe_launch_count.restype = ctypes.c_ulonglong
        _lib.b200fe_sumsq_scratch_bytes.restype = ctypes.c_size_t
    return _lib


def _vp(x):
    return ctypes.c_void_p(int(x) if x else None)


def _u(x):
    return ctypes.c_uint(int(x))


def _check(name, rc):
    if rc != 0:
        raise B200feError(name, rc)


def version():
    return lib().b200fe_version().decode()


def launch_count():
    return int(lib().b200fe_launch_count())


def last_backend():
    return lib().b200fe_last_backend().decode()


def check_device():
    return int(lib().b200fe_check_device())


def set_backend(name):
    _check("b200fe_set_backend", lib().b200fe_set_backend(name.encode()))


def set_bank_fill(mode):
    _check("b200fe_set_bank_fill", lib().b200fe_set_bank_fill(mode.encode()))


def tensor_map_available():
    return bool(lib().b200fe_tensor_map_available())


def set_gather(mode):
    """ "tma" (tiled TMA through a tensor map, default) | "cp.async": how the coa-pipe kernels fetch their tile"""
    _check("b200fe_set_gather", lib().b200fe_set_gather(mode.encode()))


QUAD_WSP = ("BwdTransQuadKernel", "BwdTransQuadKernel_Coa", "BwdTransQuadKernel_QP", "BwdTransQuadKernel_QP_1D")
QUAD_NOWSP = ("BwdTransQuadKernel_QP_Shared", "BwdTransQuadKernel_QP_1D_Shared")
HEX_WSP = ("BwdTransHexKernel", "BwdTransHexKernel_Coa", "BwdTransHexKernel_QP", "BwdTransHexKernel_QP_1D")
HEX_NOWSP = ("BwdTransHexKernel_QP_Shared", "BwdTransHexKernel_QP_1D_Shared")


def bwdtrans_quad(kernel, suf, nq0, nq1, nelmt, b0, b1, inp, out, wsp=0, stream=0, nm0=None, nm1=None, nmTot=None):
    """kernel: one of QUAD_WSP + QUAD_NOWSP; suf: 'f64' | 'f32'; b0.. are device addresses"""
    nm0 = nq0 - 1 if nm0 is None else nm0
    nm1 = nq1 - 1 if nm1 is None else nm1
    nmTot = nm0 * nm1 if nmTot is None else nmTot
    name = f"b200fe_{kernel}_{suf}"
    fn = getattr(lib(), name)
    if kernel in QUAD_NOWSP:
        rc = fn(_u(nm0), _u(nm1), _u(nmTot), _u(nq0), _u(nq1), _u(nelmt), _vp(b0), _vp(b1), _vp(inp), _vp(out),
                _vp(stream))
    else:
        rc = fn(_u(nm0), _u(nm1), _u(nmTot), _u(nq0), _u(nq1), _u(nelmt), _vp(b0), _vp(b1), _vp(inp), _vp(wsp),
                _vp(out), _vp(stream))
    _check(name, rc)


def bwdtrans_hex(kernel, suf, nq0, nq1, nq2, nelmt, b0, b1, b2, inp, out, wsp0=0, wsp1=0, stream=0, nm=None,
                 nmTot=None):
    nm0, nm1, nm2 = (nq0 - 1, nq1 - 1, nq2 - 1) if nm is None else nm
    nmTot = nm0 * nm1 * nm2 if nmTot is None else nmTot
    name = f"b200fe_{kernel}_{suf}"
    fn = getattr(lib(), name)
    head = (_u(nm0), _u(nm1), _u(nm2), _u(nmTot), _u(nq0), _u(nq1), _u(nq2), _u(nelmt), _vp(b0), _vp(b1), _vp(b2),
            _vp(inp))
    if kernel in HEX_NOWSP:
        rc = fn(*head, _vp(out), _vp(stream))
    else:
        rc = fn(*head, _vp(wsp0), _vp(wsp1), _vp(out), _vp(stream))
    _check(name, rc)


def set_data(suf, data, n, second=False, stream=0, hostgen=False):
    name = f"b200fe_set_data{'2' if second else ('_hostgen' if hostgen else '')}_{suf}"
    _check(name, getattr(lib(), name)(_vp(data), _u(n), _vp(stream)))


def l2norm_vl(suf, sums, data, n, blocks, vl, stream=0):
    name = f"b200fe_l2norm_vl_{suf}"
    _check(name, getattr(lib(), name)(_vp(sums), _vp(data), _u(n), _u(blocks), ctypes.c_int(int(vl)), _vp(stream)))


def reduce_vl(suf, sums, data, n, vl, stream=0):
    name = f"b200fe_reduce_vl_{suf}"
    _check(name, getattr(lib(), name)(_vp(sums), _vp(data), _u(n), ctypes.c_int(int(vl)), _vp(stream)))


def reduce_sum_sumsq(suf, begin, end, buffer, data, blocks, stream=0):
    name = f"b200fe_reduceSumKernel_sumsq_{suf}"
    _check(name, getattr(lib(), name)(_u(begin), _u(end), _vp(buffer), _vp(data), _u(blocks), _vp(stream)))


def add_vector(suf, x, y, n, vl, stream=0):
    name = f"b200fe_add_vector_{suf}"
    _check(name, getattr(lib(), name)(_vp(x), _vp(y), _u(n), ctypes.c_int(int(vl)), _vp(stream)))


def vector_kernel_add(suf, begin, end, x, y, stream=0):
    name = f"b200fe_vector_kernel_add_{suf}"
    _check(name, getattr(lib(), name)(_u(begin), _u(end), _vp(x), _vp(y), _vp(stream)))


def compute_matvec(suf, N, M, A, x, y, vl, stream=0):
    name = f"b200fe_compute_matvec_{suf}"
    _check(name, getattr(lib(), name)(_u(N), _u(M), _vp(A), _vp(x), _vp(y), ctypes.c_int(int(vl)), _vp(stream)))


def gemm_bwdtrans(suf, nq, nelmt, bases, inp, out, wsp, stream=0, nm=None):
    """BwdTrans in its GEMM formulation (b200fe_gemm_bwdtrans_*); nq: tuple of 2 or 3; wsp: 1 (quad) or 2 (hex) DEVICE
    scratch addresses"""
    nm = [n - 1 for n in nq] if nm is None else list(nm)
    if len(nq) == 2:
        name = f"b200fe_gemm_bwdtrans_quad_{suf}"
        rc = getattr(lib(), name)(_u(nm[0]), _u(nm[1]), _u(nq[0]), _u(nq[1]), _u(nelmt), _vp(bases[0]), _vp(bases[1]),
                                  _vp(inp), _vp(wsp[0]), _vp(out), _vp(stream))
    else:
        name = f"b200fe_gemm_bwdtrans_hex_{suf}"
        rc = getattr(lib(), name)(_u(nm[0]), _u(nm[1]), _u(nm[2]), _u(nq[0]), _u(nq[1]), _u(nq[2]), _u(nelmt),
                                  _vp(bases[0]), _vp(bases[1]), _vp(bases[2]), _vp(inp), _vp(wsp[0]), _vp(wsp[1]), _vp(out),
                                  _vp(stream))
    _check(name, rc)


def matvec_batched(suf, M, N, batch, A, strideA, x, stridex, y, stridey, stream=0):
    name = f"b200fe_matvec_batched_{suf}"
    z = ctypes.c_size_t
    _check(name, getattr(lib(), name)(_u(M), _u(N), z(int(batch)), _vp(A), z(int(strideA)), _vp(x), z(int(stridex)), _vp(y),
                                      z(int(stridey)), _vp(stream)))


def sumsq_scratch_bytes():
    return int(lib().b200fe_sumsq_scratch_bytes())


def sumsq(suf, x, n, result, scratch, stream=0):
    name = f"b200fe_sumsq_{suf}"
    _check(name, getattr(lib(), name)(_vp(x), ctypes.c_size_t(int(n)), _vp(result), _vp(scratch), _vp(stream)))


def iproduct(suf, nq, nelmt, bases, inp, out, weights=0, stream=0):
    """IProductWRTBase; nq: tuple of 2 or 3 (equal entries); DEVICE pointers; weights may be 0 (none)"""
    nm = [n - 1 for n in nq]
    if len(nq) == 2:
        name = f"b200fe_IProductWRTBaseQuad_{suf}"
        rc = getattr(lib(), name)(_u(nm[0]), _u(nm[1]), _u(nq[0]), _u(nq[1]), _u(nelmt), _vp(bases[0]), _vp(bases[1]),
                                  _vp(weights), _vp(inp), _vp(out), _vp(stream))
    else:
        name = f"b200fe_IProductWRTBaseHex_{suf}"
        rc = getattr(lib(), name)(_u(nm[0]), _u(nm[1]), _u(nm[2]), _u(nq[0]), _u(nq[1]), _u(nq[2]), _u(nelmt),
                                  _vp(bases[0]), _vp(bases[1]), _vp(bases[2]), _vp(weights), _vp(inp), _vp(out),
                                  _vp(stream))
    _check(name, rc)


def bwdtrans_sumsq(suf, nq, nelmt, bases, inp, out, sumsq, scratch, stream=0):
    """operator + checksum in one call; nq: tuple of 2 or 3; all addresses are DEVICE pointers"""
    if len(nq) == 2:
        name = f"b200fe_bwdtrans_quad_sumsq_{suf}"
        rc = getattr(lib(), name)(_u(nq[0]), _u(nq[1]), _u(nelmt), _vp(bases[0]), _vp(bases[1]), _vp(inp), _vp(out),
                                  _vp(sumsq), _vp(scratch), _vp(stream))
    else:
        name = f"b200fe_bwdtrans_hex_sumsq_{suf}"
        rc = getattr(lib(), name)(_u(nq[0]), _u(nq[1]), _u(nq[2]), _u(nelmt), _vp(bases[0]), _vp(bases[1]),
                                  _vp(bases[2]), _vp(inp), _vp(out), _vp(sumsq), _vp(scratch), _vp(stream))
    _check(name, rc)


def bwdtrans_host(suf, nq, nelmt, bases_host, in_host, out_host=0):
    """host-buffer operator; nq: tuple of 2 or 3; addresses are HOST pointers.  Returns sum(out^2)."""
    res = ctypes.c_double(0.0)
    if len(nq) == 2:
        name = f"b200fe_bwdtrans_quad_host_{suf}"
        rc = getattr(lib(), name)(_u(nq[0]), _u(nq[1]), ctypes.c_size_t(int(nelmt)), _vp(bases_host[0]),
                                  _vp(bases_host[1]), _vp(in_host), _vp(out_host), ctypes.byref(res))
    else:
        name = f"b200fe_bwdtrans_hex_host_{suf}"
        rc = getattr(lib(), name)(_u(nq[0]), _u(nq[1]), _u(nq[2]), ctypes.c_size_t(int(nelmt)),
                                  _vp(bases_host[0]), _vp(bases_host[1]), _vp(bases_host[2]), _vp(in_host),
                                  _vp(out_host), ctypes.byref(res))
    _check(name, rc)
    return res.value


class Plan:
    """b200fe_plan_*: the basis matrices uploaded once for many operator calls (include/b200fe.h).
    dim 2 = quad, 3 = hex; bases: DEVICE addresses of the dim matrices (copied by the plan)."""

    def __init__(self, dim, suf, nq, bases, stream=0):
        self.dim, self.suf, self.nq = dim, suf, nq
        self._h = ctypes.c_void_p(None)
        b = list(bases) + [0] * (3 - len(bases))
        _check("b200fe_plan_create",
               lib().b200fe_plan_create(ctypes.byref(self._h), ctypes.c_int(dim), ctypes.c_int(suf == "f32"), _u(nq),
                                        _vp(b[0]), _vp(b[1]), _vp(b[2]), _vp(stream)))

    def bwdtrans(self, nelmt, inp, out, coa=False, stream=0):
        _check("b200fe_plan_bwdtrans",
               lib().b200fe_plan_bwdtrans(self._h, ctypes.c_int(bool(coa)), _u(nelmt), _vp(inp), _vp(out),
                                          _vp(stream)))

    def iproduct(self, nelmt, inp, out, weights=0, stream=0):
        _check("b200fe_plan_iproduct",
               lib().b200fe_plan_iproduct(self._h, _u(nelmt), _vp(weights), _vp(inp), _vp(out), _vp(stream)))

    def destroy(self):
        if self._h:
            h, self._h = self._h, ctypes.c_void_p(None)
            _check("b200fe_plan_destroy", lib().b200fe_plan_destroy(h))

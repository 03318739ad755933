// ref_wrap_bwd.cu -- C launch wrappers around the REFERENCE's BwdTrans kernels
// (oracle/_ref/b04_kernels.inc, b05_kernels.inc: cut from /root/reference at build time by
// oracle/ref_build.sh).  TEST INFRASTRUCTURE: lets the tests run the reference kernels on the
// B200 next to libb200fe, with the launch shapes of the reference's run_test()
// (benchmark04.cc:907-1012, benchmark05.cc:1260-1374).  Compiled once per dtype
// (-DREF_T=double -DREF_SUF=f64 / float, f32).
#include <algorithm>
#include <chrono>
#include <type_traits>

#include <cuda_runtime.h>

#include "utils/cuda_vectors.h" // the reference's own header (found through -I<reference root>)

namespace ref04
{
#include "b04_kernels.inc"
}
namespace ref05
{
#include "b05_kernels.inc"
}

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, REF_SUF)
using T = REF_T;

// variant: 0 Uncoales, 1 Coales, 2 QP, 3 QP/Shared, 4 QP-1D, 5 QP-1D/Shared
extern "C" int FN(ref_bwdtrans_quad)(int variant, unsigned nq0, unsigned nq1, unsigned nelmt, T *basis0, T *basis1,
                                     const T *in, T *wsp0, T *wsp1, T *out, unsigned threads, unsigned elblocks,
                                     cudaStream_t s)
{
    const unsigned nm0 = nq0 - 1u, nm1 = nq1 - 1u, nmTot = nm0 * nm1;
    const unsigned blocks = nelmt / elblocks;
    const unsigned sbasis = nm0 * nq0 + nm1 * nq1;
    const unsigned sall   = sbasis + nm0 * nm1 + nq0 * nm1;
    const dim3 b2(std::min(nq0, 16u), std::min(nq1, 16u));
    switch (variant)
    {
    case 0:
        ref04::BwdTransQuadKernel<T, true><<<(nelmt + threads - 1u) / threads, threads, sizeof(T) * sbasis, s>>>(
            nm0, nm1, nmTot, nq0, nq1, nelmt, basis0, basis1, in, wsp0, out);
        break;
    case 1:
        ref04::BwdTransQuadKernel_Coa<T, true><<<(nelmt + threads - 1u) / threads, threads, sizeof(T) * sbasis, s>>>(
            nm0, nm1, nmTot, nq0, nq1, nelmt, basis0, basis1, in, wsp0, out);
        break;
    case 2:
        ref04::BwdTransQuadKernel_QP<T><<<blocks, b2, 0, s>>>(nm0, nm1, nmTot, nq0, nq1, nelmt, basis0, basis1, in,
                                                             wsp1, out);
        break;
    case 3:
        ref04::BwdTransQuadKernel_QP<T><<<blocks, b2, sizeof(T) * sall, s>>>(nm0, nm1, nmTot, nq0, nq1, nelmt, basis0,
                                                                            basis1, in, out);
        break;
    case 4:
        ref04::BwdTransQuadKernel_QP_1D<T><<<blocks, std::min(nq0 * nq1, threads), 0, s>>>(
            nm0, nm1, nmTot, nq0, nq1, nelmt, basis0, basis1, in, wsp1, out);
        break;
    case 5:
        ref04::BwdTransQuadKernel_QP_1D<T><<<blocks, std::min(nq0 * nq1, threads), sizeof(T) * sall, s>>>(
            nm0, nm1, nmTot, nq0, nq1, nelmt, basis0, basis1, in, out);
        break;
    default:
        return -1;
    }
    return (int)cudaGetLastError();
}

// wsp0: nelmt*nm1*nm2, wsp1: nelmt*nm2 (thread-per-element variants);
// wsp2: nelmt*nq0*nm1*nm2, wsp3: nelmt*nq0*nq1*nm2 (QP variants)
extern "C" int FN(ref_bwdtrans_hex)(int variant, unsigned nq0, unsigned nq1, unsigned nq2, unsigned nelmt, T *basis0,
                                    T *basis1, T *basis2, const T *in, T *wsp0, T *wsp1, T *wsp2, T *wsp3, T *out,
                                    unsigned threads, unsigned elblocks, cudaStream_t s)
{
    const unsigned nm0 = nq0 - 1u, nm1 = nq1 - 1u, nm2 = nq2 - 1u, nmTot = nm0 * nm1 * nm2;
    const unsigned blocks = nelmt / elblocks;
    const unsigned sbasis = nm0 * nq0 + nm1 * nq1 + nm2 * nq2;
    const unsigned sall   = sbasis + nm0 * nm1 * nm2 + nm0 * nm1 * nq2 + nm0 * nq1 * nq2;
    const dim3 b3(std::min(nq0, 16u), std::min(nq1, 16u), std::min(nq2, std::max(1u, 254u / (nq0 * nq1))));
    switch (variant)
    {
    case 0:
        ref05::BwdTransHexKernel<T, true><<<(nelmt + threads - 1u) / threads, threads, sizeof(T) * sbasis, s>>>(
            nm0, nm1, nm2, nmTot, nq0, nq1, nq2, nelmt, basis0, basis1, basis2, in, wsp0, wsp1, out);
        break;
    case 1:
        ref05::BwdTransHexKernel_Coa<T, true><<<(nelmt + threads - 1u) / threads, threads, sizeof(T) * sbasis, s>>>(
            nm0, nm1, nm2, nmTot, nq0, nq1, nq2, nelmt, basis0, basis1, basis2, in, wsp0, wsp1, out);
        break;
    case 2:
        ref05::BwdTransHexKernel_QP<T><<<blocks, b3, 0, s>>>(nm0, nm1, nm2, nmTot, nq0, nq1, nq2, nelmt, basis0, basis1,
                                                            basis2, in, wsp2, wsp3, out);
        break;
    case 3:
        ref05::BwdTransHexKernel_QP<T><<<blocks, b3, sizeof(T) * sall, s>>>(nm0, nm1, nm2, nmTot, nq0, nq1, nq2, nelmt,
                                                                           basis0, basis1, basis2, in, out);
        break;
    case 4:
        ref05::BwdTransHexKernel_QP_1D<T><<<blocks, std::min(nq0 * nq1 * nq2, threads), 0, s>>>(
            nm0, nm1, nm2, nmTot, nq0, nq1, nq2, nelmt, basis0, basis1, basis2, in, wsp2, wsp3, out);
        break;
    case 5:
        ref05::BwdTransHexKernel_QP_1D<T><<<blocks, std::min(nq0 * nq1 * nq2, threads), sizeof(T) * sall, s>>>(
            nm0, nm1, nm2, nmTot, nq0, nq1, nq2, nelmt, basis0, basis1, basis2, in, out);
        break;
    default:
        return -1;
    }
    return (int)cudaGetLastError();
}

// sumfac_generic.cuh -- "generic" back-end: any nq0 x nq1 (x nq2), sizes known
// only at run time.  Safety net for the shapes the tuned back-ends do not
// instantiate (unequal or very large nq); same arithmetic and summation order
// as the tuned kernels, basis and all intermediates in shared memory.
//
// A CTA walks elements e = blockIdx.x, blockIdx.x + gridDim.x, ...; the threads
// split each contraction pass over its flattened output index.  `coa` selects
// the warp-interleaved global layout.
#pragma once

#include "common.cuh"

namespace b200fe
{

__device__ __forceinline__ size_t gidx(bool coa, size_t e, unsigned k, unsigned len)
{
    return coa ? (e >> 5) * 32u * (size_t)len + 32u * (size_t)k + (e & 31u) : e * (size_t)len + k;
}

template <typename T>
__global__ void bwdtrans_quad_generic_kernel(unsigned nm0, unsigned nm1, unsigned nq0, unsigned nq1, unsigned nelmt,
                                             const T *__restrict__ basis0, const T *__restrict__ basis1,
                                             const T *__restrict__ in, T *__restrict__ out, int coa_flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sb0 = reinterpret_cast<T *>(smem_raw); // [p][i]
    T *sb1 = sb0 + nm0 * nq0;                 // [q][j]
    T *s0  = sb1 + nm1 * nq1;                 // in   [q][p]
    T *s1  = s0 + nm0 * nm1;                  // dir0 [i][q]
    const bool coa       = coa_flag != 0;
    const unsigned nmTot = nm0 * nm1, nqTot = nq0 * nq1;

    for (unsigned t = threadIdx.x; t < nm0 * nq0; t += blockDim.x)
        sb0[t] = basis0[t];
    for (unsigned t = threadIdx.x; t < nm1 * nq1; t += blockDim.x)
        sb1[t] = basis1[t];

    for (size_t e = blockIdx.x; e < nelmt; e += gridDim.x)
    {
        for (unsigned t = threadIdx.x; t < nmTot; t += blockDim.x)
            s0[t] = in[gidx(coa, e, t, nmTot)];
        __syncthreads();
        for (unsigned t = threadIdx.x; t < nq0 * nm1; t += blockDim.x)
        {
            const unsigned i = t / nm1, q = t - i * nm1;
            T acc = T(0);
            for (unsigned p = 0; p < nm0; ++p)
                acc = fmadd(s0[q * nm0 + p], sb0[p * nq0 + i], acc);
            s1[t] = acc;
        }
        __syncthreads();
        for (unsigned t = threadIdx.x; t < nqTot; t += blockDim.x)
        {
            const unsigned j = t / nq0, i = t - j * nq0;
            T acc = T(0);
            for (unsigned q = 0; q < nm1; ++q)
                acc = fmadd(s1[i * nm1 + q], sb1[q * nq1 + j], acc);
            out[gidx(coa, e, t, nqTot)] = acc;
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void bwdtrans_hex_generic_kernel(unsigned nm0, unsigned nm1, unsigned nm2, unsigned nq0, unsigned nq1,
                                            unsigned nq2, unsigned nelmt, const T *__restrict__ basis0,
                                            const T *__restrict__ basis1, const T *__restrict__ basis2,
                                            const T *__restrict__ in, T *__restrict__ out, int coa_flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sb0 = reinterpret_cast<T *>(smem_raw);
    T *sb1 = sb0 + nm0 * nq0;
    T *sb2 = sb1 + nm1 * nq1;
    T *s0  = sb2 + nm2 * nq2;       // in   [r][q][p]
    T *s1  = s0 + nm0 * nm1 * nm2;  // dir0 [i][r][q]
    T *s2  = s1 + nq0 * nm1 * nm2;  // dir1 [j][i][r]
    const bool coa       = coa_flag != 0;
    const unsigned nmTot = nm0 * nm1 * nm2, nqTot = nq0 * nq1 * nq2;

    for (unsigned t = threadIdx.x; t < nm0 * nq0; t += blockDim.x)
        sb0[t] = basis0[t];
    for (unsigned t = threadIdx.x; t < nm1 * nq1; t += blockDim.x)
        sb1[t] = basis1[t];
    for (unsigned t = threadIdx.x; t < nm2 * nq2; t += blockDim.x)
        sb2[t] = basis2[t];

    for (size_t e = blockIdx.x; e < nelmt; e += gridDim.x)
    {
        for (unsigned t = threadIdx.x; t < nmTot; t += blockDim.x)
            s0[t] = in[gidx(coa, e, t, nmTot)];
        __syncthreads();
        for (unsigned t = threadIdx.x; t < nq0 * nm1 * nm2; t += blockDim.x)
        {
            const unsigned i = t / (nm1 * nm2), rq = t - i * (nm1 * nm2);
            T acc = T(0);
            for (unsigned p = 0; p < nm0; ++p)
                acc = fmadd(s0[rq * nm0 + p], sb0[p * nq0 + i], acc);
            s1[t] = acc;
        }
        __syncthreads();
        for (unsigned t = threadIdx.x; t < nq1 * nq0 * nm2; t += blockDim.x)
        {
            const unsigned j = t / (nq0 * nm2), ir = t - j * (nq0 * nm2);
            const unsigned i = ir / nm2, r = ir - i * nm2;
            T acc = T(0);
            for (unsigned q = 0; q < nm1; ++q)
                acc = fmadd(s1[(i * nm2 + r) * nm1 + q], sb1[q * nq1 + j], acc);
            s2[t] = acc;
        }
        __syncthreads();
        for (unsigned t = threadIdx.x; t < nqTot; t += blockDim.x)
        {
            const unsigned k = t / (nq0 * nq1), ji = t - k * (nq0 * nq1);
            T acc = T(0);
            for (unsigned r = 0; r < nm2; ++r)
                acc = fmadd(s2[ji * nm2 + r], sb2[r * nq2 + k], acc);
            out[gidx(coa, e, t, nqTot)] = acc;
        }
        __syncthreads();
    }
}

template <typename T>
inline int launch_quad_generic(unsigned nm0, unsigned nm1, unsigned nq0, unsigned nq1, unsigned nelmt, const T *b0,
                               const T *b1, const T *in, T *out, bool coa, cudaStream_t stream)
{
    // 64-bit products: caller-supplied shapes must not wrap into something that passes the check
    if (nm0 > 0xffffu || nm1 > 0xffffu || nq0 > 0xffffu || nq1 > 0xffffu)
        return B200FE_EUNSUPPORTED;
    const size_t smem = ((size_t)nm0 * nq0 + (size_t)nm1 * nq1 + (size_t)nm0 * nm1 + (size_t)nq0 * nm1) * sizeof(T);
    if (smem > (size_t)kSmemMax)
        return B200FE_EUNSUPPORTED;
    B200FE_CUDA_TRY(cudaFuncSetAttribute(bwdtrans_quad_generic_kernel<T>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned threads = nq0 * nq1;
    threads          = threads < 64 ? 64 : (threads > 256 ? 256 : (threads + 31) / 32 * 32);
    const unsigned grid = nelmt < 148u * 16u ? nelmt : 148u * 16u;
    bwdtrans_quad_generic_kernel<T>
        <<<grid, threads, smem, stream>>>(nm0, nm1, nq0, nq1, nelmt, b0, b1, in, out, coa ? 1 : 0);
    count_launch();
    return launch_status();
}

template <typename T>
inline int launch_hex_generic(unsigned nm0, unsigned nm1, unsigned nm2, unsigned nq0, unsigned nq1, unsigned nq2,
                              unsigned nelmt, const T *b0, const T *b1, const T *b2, const T *in, T *out, bool coa,
                              cudaStream_t stream)
{
    if (nm0 > 0x3ffu || nm1 > 0x3ffu || nm2 > 0x3ffu || nq0 > 0x3ffu || nq1 > 0x3ffu || nq2 > 0x3ffu)
        return B200FE_EUNSUPPORTED; // below 2^10 each: no 64-bit product of three can wrap, and nothing larger fits
    const size_t smem = ((size_t)nm0 * nq0 + (size_t)nm1 * nq1 + (size_t)nm2 * nq2 + (size_t)nm0 * nm1 * nm2 +
                         (size_t)nq0 * nm1 * nm2 + (size_t)nq0 * nq1 * nm2) *
                        sizeof(T);
    if (smem > (size_t)kSmemMax)
        return B200FE_EUNSUPPORTED;
    B200FE_CUDA_TRY(cudaFuncSetAttribute(bwdtrans_hex_generic_kernel<T>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned threads = nq0 * nq1 * nq2;
    threads          = threads < 64 ? 64 : (threads > 256 ? 256 : (threads + 31) / 32 * 32);
    const unsigned grid = nelmt < 148u * 8u ? nelmt : 148u * 8u;
    bwdtrans_hex_generic_kernel<T>
        <<<grid, threads, smem, stream>>>(nm0, nm1, nm2, nq0, nq1, nq2, nelmt, b0, b1, b2, in, out, coa ? 1 : 0);
    count_launch();
    return launch_status();
}

} // namespace b200fe

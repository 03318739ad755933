// sumfac_tpe.cuh -- "tpe" back-end: one thread per element, whole element in
// registers, for the warp-interleaved layout x[(e/32)*32*len + 32*k + e%32]
// (reference BwdTrans*Kernel_Coa, benchmark04.cc:78-147, benchmark05.cc:104-201).
//
// In that layout lane l of a warp owns element 32*g + l and, for every
// in-element index k, the warp touches 32 consecutive values: all global
// accesses are full-line coalesced with no staging at all.  The element's
// modes live in registers for the whole computation (read from HBM exactly
// once -- the reference re-reads them nq0 times through L1/L2), intermediates
// never leave the register file (the reference round-trips them through a
// global wsp array), and the basis comes from the constant bank.
//
// Loop nest and summation order are the reference's: for every i, contract
// direction 0, then for every j direction 1 (then for every k direction 2).
#pragma once

#include "common.cuh"

namespace b200fe
{

// VW consecutive elements (lanes of an interleave group) per thread, fetched / stored as ONE vector: the smallest
// operators (nq = 2: one value in, four out per element) otherwise have 4 bytes per thread in flight and are
// latency-bound (FP32 quad nq = 2: 0.70 of the roofline with one element per thread).  Per element the arithmetic and
// its order are unchanged.
template <typename T, int VW> struct alignas(sizeof(T) * VW) TpeVec
{
    T v[VW];
};
template <typename T, int VW> __device__ __forceinline__ void tpe_load(T (&dst)[VW], const T *p)
{
    if constexpr (VW == 1)
        dst[0] = ld_stream(p);
    else if constexpr (sizeof(T) * VW == 16)
    {
        const float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
        *reinterpret_cast<float4 *>(dst) = t;
    }
    else
    {
        const float2 t = __ldcs(reinterpret_cast<const float2 *>(p));
        *reinterpret_cast<float2 *>(dst) = t;
    }
}
template <typename T, int VW> __device__ __forceinline__ void tpe_store(T *p, const T (&src)[VW])
{
    if constexpr (VW == 1)
        st_stream(p, src[0]);
    else if constexpr (sizeof(T) * VW == 16)
        __stcs(reinterpret_cast<float4 *>(p), *reinterpret_cast<const float4 *>(src));
    else
        __stcs(reinterpret_cast<float2 *>(p), *reinterpret_cast<const float2 *>(src));
}

template <typename T, int NQ, int THREADS, int VW>
__device__ __noinline__ void bwdtrans_quad_tpe_coa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int THREADS, int VW = 1>
__global__ void __launch_bounds__(THREADS)
    bwdtrans_quad_tpe_coa_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    bwdtrans_quad_tpe_coa_body<T, NQ, THREADS, VW>(in, out, nelmt);
}

template <typename T, int NQ, int THREADS, int VW>
__device__ __noinline__ void bwdtrans_quad_tpe_coa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    constexpr int NM = NQ - 1, NM2 = NM * NM, NQ2 = NQ * NQ;
    constexpr int BP = bank_pitch<T>(NQ), B0 = 0, B1 = NM * BP; // pitched bank rows (common.cuh)
    const size_t e = ((size_t)blockIdx.x * THREADS + threadIdx.x) * VW; // first of this thread's VW elements
    if (e >= nelmt) // (nelmt % 32 == 0: whole vectors)
        return;
    const size_t g = e >> 5, l = e & 31;
    const T *pin = in + g * 32 * NM2 + l;
    T *pout      = out + g * 32 * NQ2 + l;

    alignas(16) T a[NM2][VW];
#pragma unroll
    for (int k = 0; k < NM2; ++k)
        tpe_load<T, VW>(a[k], pin + 32 * k);

#pragma unroll
    for (int i = 0; i < NQ; ++i)
    {
        T w[NM][VW];
#pragma unroll
        for (int q = 0; q < NM; ++q)
#pragma unroll
            for (int v = 0; v < VW; ++v)
            {
                T t = T(0);
#pragma unroll
                for (int p = 0; p < NM; ++p)
                    t = fmadd(a[q * NM + p][v], cbasis<T>(B0 + p * BP + i), t);
                w[q][v] = t;
            }
#pragma unroll
        for (int j = 0; j < NQ; ++j)
        {
            alignas(16) T o[VW];
#pragma unroll
            for (int v = 0; v < VW; ++v)
            {
                T t = T(0);
#pragma unroll
                for (int q = 0; q < NM; ++q)
                    t = fmadd(w[q][v], cbasis<T>(B1 + q * BP + j), t);
                o[v] = t;
            }
            tpe_store<T, VW>(pout + 32 * (j * NQ + i), o);
        }
    }
}

template <typename T, int NQ, int THREADS, int VW>
__device__ __noinline__ void bwdtrans_hex_tpe_coa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int THREADS>
__device__ __noinline__ void bwdtrans_hex_tpe_coa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int THREADS, int VW = 1>
__global__ void __launch_bounds__(THREADS)
    bwdtrans_hex_tpe_coa_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    bwdtrans_hex_tpe_coa_body<T, NQ, THREADS, VW>(in, out, nelmt);
}

template <typename T, int NQ, int THREADS, int VW>
__device__ __noinline__ void bwdtrans_hex_tpe_coa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    constexpr int NM = NQ - 1, NM2 = NM * NM, NM3 = NM2 * NM, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
    constexpr int BP = bank_pitch<T>(NQ), B0 = 0, B1 = NM * BP, B2 = 2 * NM * BP;
    const size_t e = ((size_t)blockIdx.x * THREADS + threadIdx.x) * VW;
    if (e >= nelmt)
        return;
    const size_t g = e >> 5, l = e & 31;
    const T *pin = in + g * 32 * NM3 + l;
    T *pout      = out + g * 32 * NQ3 + l; // intended offset (reference bug at benchmark05.cc:193 not reproduced)

    alignas(16) T a[NM3][VW];
#pragma unroll
    for (int k = 0; k < NM3; ++k)
        tpe_load<T, VW>(a[k], pin + 32 * k);

#pragma unroll
    for (int i = 0; i < NQ; ++i)
    {
        T w0[NM2][VW];
#pragma unroll
        for (int rq = 0; rq < NM2; ++rq)
#pragma unroll
            for (int v = 0; v < VW; ++v)
            {
                T t = T(0);
#pragma unroll
                for (int p = 0; p < NM; ++p)
                    t = fmadd(a[rq * NM + p][v], cbasis<T>(B0 + p * BP + i), t);
                w0[rq][v] = t;
            }
#pragma unroll
        for (int j = 0; j < NQ; ++j)
        {
            T w1[NM][VW];
#pragma unroll
            for (int r = 0; r < NM; ++r)
#pragma unroll
                for (int v = 0; v < VW; ++v)
                {
                    T t = T(0);
#pragma unroll
                    for (int q = 0; q < NM; ++q)
                        t = fmadd(w0[r * NM + q][v], cbasis<T>(B1 + q * BP + j), t);
                    w1[r][v] = t;
                }
#pragma unroll
            for (int k = 0; k < NQ; ++k)
            {
                alignas(16) T o[VW];
#pragma unroll
                for (int v = 0; v < VW; ++v)
                {
                    T t = T(0);
#pragma unroll
                    for (int r = 0; r < NM; ++r)
                        t = fmadd(w1[r][v], cbasis<T>(B2 + r * BP + k), t);
                    o[v] = t;
                }
                tpe_store<T, VW>(pout + 32 * (k * NQ2 + j * NQ + i), o);
            }
        }
    }
}

} // namespace b200fe

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

template <typename T, int NQ, int THREADS>
__device__ __noinline__ void bwdtrans_quad_tpe_coa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int THREADS>
__global__ void __launch_bounds__(THREADS)
    bwdtrans_quad_tpe_coa_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    bwdtrans_quad_tpe_coa_body<T, NQ, THREADS>(in, out, nelmt);
}

template <typename T, int NQ, int THREADS>
__device__ __noinline__ void bwdtrans_quad_tpe_coa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    constexpr int NM = NQ - 1, NM2 = NM * NM, NQ2 = NQ * NQ;
    constexpr int BP = bank_pitch<T>(NQ), B0 = 0, B1 = NM * BP; // pitched bank rows (common.cuh)
    const size_t e = (size_t)blockIdx.x * THREADS + threadIdx.x;
    if (e >= nelmt)
        return;
    const size_t g = e >> 5, l = e & 31;
    const T *pin = in + g * 32 * NM2 + l;
    T *pout      = out + g * 32 * NQ2 + l;

    T a[NM2];
#pragma unroll
    for (int k = 0; k < NM2; ++k)
        a[k] = ld_stream(pin + 32 * k);

#pragma unroll
    for (int i = 0; i < NQ; ++i)
    {
        T w[NM];
#pragma unroll
        for (int q = 0; q < NM; ++q)
        {
            T t = T(0);
#pragma unroll
            for (int p = 0; p < NM; ++p)
                t = fmadd(a[q * NM + p], cbasis<T>(B0 + p * BP + i), t);
            w[q] = t;
        }
#pragma unroll
        for (int j = 0; j < NQ; ++j)
        {
            T t = T(0);
#pragma unroll
            for (int q = 0; q < NM; ++q)
                t = fmadd(w[q], cbasis<T>(B1 + q * BP + j), t);
            st_stream(pout + 32 * (j * NQ + i), t);
        }
    }
}

template <typename T, int NQ, int THREADS>
__device__ __noinline__ void bwdtrans_hex_tpe_coa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int THREADS>
__global__ void __launch_bounds__(THREADS)
    bwdtrans_hex_tpe_coa_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    bwdtrans_hex_tpe_coa_body<T, NQ, THREADS>(in, out, nelmt);
}

template <typename T, int NQ, int THREADS>
__device__ __noinline__ void bwdtrans_hex_tpe_coa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    constexpr int NM = NQ - 1, NM2 = NM * NM, NM3 = NM2 * NM, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
    constexpr int BP = bank_pitch<T>(NQ), B0 = 0, B1 = NM * BP, B2 = 2 * NM * BP;
    const size_t e = (size_t)blockIdx.x * THREADS + threadIdx.x;
    if (e >= nelmt)
        return;
    const size_t g = e >> 5, l = e & 31;
    const T *pin = in + g * 32 * NM3 + l;
    T *pout      = out + g * 32 * NQ3 + l; // intended offset (reference bug at benchmark05.cc:193 not reproduced)

    T a[NM3];
#pragma unroll
    for (int k = 0; k < NM3; ++k)
        a[k] = ld_stream(pin + 32 * k);

#pragma unroll
    for (int i = 0; i < NQ; ++i)
    {
        T w0[NM2];
#pragma unroll
        for (int rq = 0; rq < NM2; ++rq)
        {
            T t = T(0);
#pragma unroll
            for (int p = 0; p < NM; ++p)
                t = fmadd(a[rq * NM + p], cbasis<T>(B0 + p * BP + i), t);
            w0[rq] = t;
        }
#pragma unroll
        for (int j = 0; j < NQ; ++j)
        {
            T w1[NM];
#pragma unroll
            for (int r = 0; r < NM; ++r)
            {
                T t = T(0);
#pragma unroll
                for (int q = 0; q < NM; ++q)
                    t = fmadd(w0[r * NM + q], cbasis<T>(B1 + q * BP + j), t);
                w1[r] = t;
            }
#pragma unroll
            for (int k = 0; k < NQ; ++k)
            {
                T t = T(0);
#pragma unroll
                for (int r = 0; r < NM; ++r)
                    t = fmadd(w1[r], cbasis<T>(B2 + r * BP + k), t);
                st_stream(pout + 32 * (k * NQ2 + j * NQ + i), t);
            }
        }
    }
}

} // namespace b200fe

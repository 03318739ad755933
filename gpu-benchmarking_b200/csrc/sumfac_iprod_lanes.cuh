// sumfac_iprod_lanes.cuh -- IProductWRTBase in the style of sumfac_lanes.cuh: one row (or plane) per
// thread and phase, every index a compile-time constant, the basis operand from the TRANSPOSED constant bank
// (common.cuh: bank[(d*nq + i)*pitch(nm) + p] = B_d[p*nq + i]).
//   quad  out[e][q][p]    = sum_j B1[q][j] ( sum_i B0[p][i] (w*in)[e][j][i] )
//   hex   out[e][r][q][p] = sum_k B2[r][k] ( sum_j B1[q][j] ( sum_i B0[p][i] (w*in)[e][k][j][i] ) )
// A row of the input is nq CONTIGUOUS values (a plane nq^2), an even count for the even nq served here, so a thread
// fetches its own row / plane straight from global memory with 8- or 16-byte loads (consecutive threads take
// consecutive rows: a warp covers one contiguous stretch, every sector is used in full, the second half of a sector
// hits L1) -- no staging, no padded tile, no barrier before direction 0.  w*in is one rounded product, every sum runs
// in ascending index order from +0 with fused multiply-adds: bit-identical to oracle_iproduct_* and to
// sumfac_iprod.cuh.
//   quad  phase A  thread (e, j): row -> t[j][p] to shared memory as [j][e][p] (row stride = 1 mod 32)
//         phase B  thread (e, p), p fastest: column t[.][p] -> out[e][.][p] straight to global
//   hex   phase A  thread (e, k): plane -> directions 0 and 1 in registers -> u[k][(q, p)] to shared memory
//         phase B  threads over (e, (q, p)), (q, p) fastest: u[.][(q, p)] -> out[e][.][q][p], runs of nm^2 values
#pragma once

#include <type_traits>

#include "sumfac_lanes.cuh"

namespace b200fe
{

// N contiguous values at g (16-byte aligned when VEC16, else 8-byte aligned; N even), times the metric where there is one
template <typename T, int N, bool VEC16, bool WEIGHTED>
__device__ __forceinline__ void iprod_fetch(T (&x)[N], const T *__restrict__ g, const T *__restrict__ w)
{
    constexpr int W = VEC16 ? 16 / (int)sizeof(T) : 8 / (int)sizeof(T); // values per load
    static_assert(N % W == 0, "whole vectors");
    using V = typename std::conditional<VEC16, typename Vec16<T>::type,
                                        typename std::conditional<sizeof(T) == 4, float2, double>::type>::type;
    V v[N / W], u[N / W];
#pragma unroll
    for (int c = 0; c < N / W; ++c)
        v[c] = __ldg(reinterpret_cast<const V *>(g) + c); // through L1: the two halves of a sector are separate loads
    if (WEIGHTED)
    {
#pragma unroll
        for (int c = 0; c < N / W; ++c)
            u[c] = __ldg(reinterpret_cast<const V *>(w) + c);
    }
#pragma unroll
    for (int c = 0; c < N / W; ++c)
    {
        const T *pv = reinterpret_cast<const T *>(&v[c]);
        const T *pu = reinterpret_cast<const T *>(&u[c]);
#pragma unroll
        for (int k = 0; k < W; ++k)
            x[c * W + k] = WEIGHTED ? pv[k] * pu[k] : pv[k];
    }
}

template <typename T, int NQ, int EL> struct QuadIprodLanes
{
    static_assert(NQ % 2 == 0, "rows must be whole 8-byte vectors");
    static constexpr int NM = NQ - 1, NQ2 = NQ * NQ, NM2 = NM * NM;
    static constexpr int THREADS = EL * NQ;
    static constexpr int J1      = (EL * NM + 30) / 32 * 32 + 1; // row stride of t: = 1 mod 32
    static constexpr size_t SMEM = (size_t)NQ * J1 * sizeof(T);
    static constexpr bool VEC16  = (NQ * sizeof(T)) % 16 == 0;
    static constexpr int B0 = 0, B1 = NQ * bank_pitch<T>(NM);
    static_assert(THREADS <= 1024, "block size");
};

template <typename T, int NQ, int EL, bool WEIGHTED>
__device__ __noinline__ void iproduct_quad_lanes_body(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int EL, bool WEIGHTED>
__global__ void __launch_bounds__(QuadIprodLanes<T, NQ, EL>::THREADS)
    iproduct_quad_lanes_kernel(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    iproduct_quad_lanes_body<T, NQ, EL, WEIGHTED>(in, w, out, nelmt);
}

template <typename T, int NQ, int EL, bool WEIGHTED>
__device__ __noinline__ void iproduct_quad_lanes_body(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out, unsigned nelmt)
{
    using C = QuadIprodLanes<T, NQ, EL>;
    constexpr int NM = C::NM, J1 = C::J1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s1 = reinterpret_cast<T *>(smem_raw);
    const int tid   = threadIdx.x;
    const size_t e0 = (size_t)blockIdx.x * EL;
    const int ne    = (nelmt - e0 < (size_t)EL) ? (int)(nelmt - e0) : EL;
    {
        const int e = tid / NQ, j = tid - e * NQ;
        T x[NQ];
        const size_t off = ((e0 + (e < ne ? e : 0)) * NQ + j) * NQ; // clamp: a ragged tile computes something harmless
        iprod_fetch<T, NQ, C::VEC16, WEIGHTED>(x, in + off, WEIGHTED ? w + off : nullptr);
        T *dst = s1 + j * J1 + e * NM;
        lanes_row<T, NQ, NM, C::B0>(x, [&](int p, T v) { dst[p] = v; });
    }
    __syncthreads();
    if (tid < EL * NM)
    {
        const int e = tid / NM, p = tid - e * NM;
        T x[NQ];
        const T *src = s1 + tid;
#pragma unroll
        for (int j = 0; j < NQ; ++j)
            x[j] = src[j * J1];
        if (e < ne)
        {
            T *dst = out + (e0 + e) * C::NM2 + p;
            lanes_row<T, NQ, NM, C::B1>(x, [&](int q, T v) { st_stream(dst + q * NM, v); });
        }
    }
}

// STAGED: the planes of a hex are 128-512 bytes apart, so the per-thread vector loads above touch a different line in
// every lane (nq = 8 FP32: 16 loads x 32 lines per warp for 64 lines of data) and the L1 tag rate, not HBM, sets the
// pace (0.64-0.71 of the roofline).  The staged variant loads the CTA's slab cooperatively (consecutive threads,
// consecutive 16-byte vectors, metric multiplied in) into shared memory with the plane stride padded by one vector
// (bank shift 4 words per plane: the threads' own 16-byte plane reads are conflict-free) and takes the planes from
// there.
template <typename T, int NQ, int EL, bool STAGED = false> struct HexIprodLanes
{
    static_assert(NQ % 2 == 0, "planes must be whole 16-byte vectors");
    static constexpr int NM = NQ - 1, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ, NM2 = NM * NM, NM3 = NM2 * NM;
    static constexpr int THREADS = (EL * NQ + 31) / 32 * 32;
    static constexpr int K1      = (EL * NM2 + 30) / 32 * 32 + 1; // plane stride of u: = 1 mod 32
    static constexpr int W       = 16 / (int)sizeof(T);
    static constexpr int PS      = NQ2 + W;                       // padded plane stride of the staged slab
    static constexpr int S2      = (NQ * K1 * (int)sizeof(T) + 15) / 16 * 16 / (int)sizeof(T);
    static constexpr size_t SMEM = (size_t)(S2 + (STAGED ? EL * NQ * PS : 0)) * sizeof(T);
    static constexpr int B2      = 2 * NQ * bank_pitch<T>(NM);
    static constexpr int IB0     = sizeof(T) == 4 ? 2 : 1;
    static constexpr bool ROLLED = sizeof(T) == 8 || NQ >= 8; // see HexLanes
};

template <typename T, int NQ, int EL, bool WEIGHTED, int MINB, bool STAGED>
__device__ __noinline__ void iproduct_hex_lanes_body(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int EL, bool WEIGHTED, int MINB = 1, bool STAGED = false>
__global__ void __launch_bounds__(HexIprodLanes<T, NQ, EL, STAGED>::THREADS, MINB)
    iproduct_hex_lanes_kernel(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    iproduct_hex_lanes_body<T, NQ, EL, WEIGHTED, MINB, STAGED>(in, w, out, nelmt);
}

template <typename T, int NQ, int EL, bool WEIGHTED, int MINB, bool STAGED>
__device__ __noinline__ void iproduct_hex_lanes_body(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out, unsigned nelmt)
{
    using C = HexIprodLanes<T, NQ, EL, STAGED>;
    constexpr int NM = C::NM, NQ2 = C::NQ2, NM2 = C::NM2, K1 = C::K1, IB0 = C::IB0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s2 = reinterpret_cast<T *>(smem_raw);
    const int tid   = threadIdx.x;
    const size_t e0 = (size_t)blockIdx.x * EL;
    const int ne    = (nelmt - e0 < (size_t)EL) ? (int)(nelmt - e0) : EL;
    if constexpr (STAGED)
    {
        using V         = typename Vec16<T>::type;
        constexpr int W = C::W, VPP = NQ2 / W; // vectors per plane
        T *s_in         = s2 + C::S2;
        const V *gin    = reinterpret_cast<const V *>(in + e0 * C::NQ3);
        const V *gw     = reinterpret_cast<const V *>(w + (WEIGHTED ? e0 * C::NQ3 : 0));
        const int nv    = ne * NQ * VPP;
        constexpr int ITER = (EL * NQ * VPP + C::THREADS - 1) / C::THREADS;
        V v[ITER], u[ITER];
#pragma unroll
        for (int it = 0; it < ITER; ++it)
        {
            const int c = it * C::THREADS + tid;
            if (c < nv)
            {
                v[it] = ld_stream(gin + c);
                if (WEIGHTED)
                    u[it] = ld_stream(gw + c);
            }
        }
#pragma unroll
        for (int it = 0; it < ITER; ++it)
        {
            const int c = it * C::THREADS + tid;
            if (c < nv)
            {
                if (WEIGHTED)
                {
                    T *pv       = reinterpret_cast<T *>(&v[it]);
                    const T *pu = reinterpret_cast<const T *>(&u[it]);
#pragma unroll
                    for (int k = 0; k < W; ++k)
                        pv[k] = pv[k] * pu[k];
                }
                const int plane = c / VPP, off = c - plane * VPP;
                *reinterpret_cast<V *>(s_in + plane * C::PS + off * W) = v[it];
            }
        }
        __syncthreads();
    }
    if (tid < EL * NQ)
    {
        const int e = tid / NQ, k = tid - e * NQ;
        T a[NQ2];
        if constexpr (STAGED)
        {
            using V      = typename Vec16<T>::type;
            const V *src = reinterpret_cast<const V *>(s2 + C::S2 + tid * C::PS);
#pragma unroll
            for (int c = 0; c < NQ2 / C::W; ++c)
            {
                const V t = src[c];
                const T *pt = reinterpret_cast<const T *>(&t);
#pragma unroll
                for (int kk = 0; kk < C::W; ++kk)
                    a[c * C::W + kk] = pt[kk];
            }
        }
        else
        {
            const size_t off = ((e0 + (e < ne ? e : 0)) * NQ + k) * NQ2;
            iprod_fetch<T, NQ2, true, WEIGHTED>(a, in + off, WEIGHTED ? w + off : nullptr);
        }
        T *dst = s2 + k * K1 + e * NM2; // u[(q, p)] at dst[q*NM + p]
        if constexpr (C::ROLLED)
        {
#pragma unroll 1
            for (int pb = 0; pb + IB0 <= NM; pb += IB0)
                plane_block<T, NQ, NM, 1, IB0>(a, dst, pb);
        }
        else
        {
#pragma unroll
            for (int pb = 0; pb + IB0 <= NM; pb += IB0)
                plane_block<T, NQ, NM, 1, IB0>(a, dst, pb);
        }
        if constexpr (NM % IB0 != 0) // nm is odd: the last output p alone
            plane_block<T, NQ, NM, 1, 1>(a, dst, NM - 1);
    }
    __syncthreads();

    constexpr int ROWS = EL * NM2, ITER = (ROWS + C::THREADS - 1) / C::THREADS;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it)
    {
        const int row = it * C::THREADS + tid;
        if (row >= ROWS)
            break;
        const int e = row / NM2, qp = row - e * NM2;
        T x[NQ];
        const T *src = s2 + row;
#pragma unroll
        for (int k = 0; k < NQ; ++k)
            x[k] = src[k * K1];
        if (e < ne)
        {
            T *dst = out + (e0 + e) * C::NM3 + qp;
            lanes_row<T, NQ, NM, C::B2>(x, [&](int r, T v) { st_stream(dst + r * NM2, v); });
        }
    }
}

} // namespace b200fe

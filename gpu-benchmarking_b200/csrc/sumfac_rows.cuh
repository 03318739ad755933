// sumfac_rows.cuh -- "rows" back-end: element-batched sum-factorisation.
//
// One CTA owns a tile of E consecutive elements (element-major layout, so the
// tile is ONE contiguous slab of global memory in and one out).  Each
// contraction direction is a pass over "rows": a thread takes one contiguous
// row of NM intermediate values out of shared memory into registers and
// produces the NQ values of that row in the new direction, every (p, i) loop
// unrolled so the basis operand is a constant-bank immediate.  Between passes
// the tile is re-laid out in shared memory so the next direction's rows are
// contiguous again; rows are NM (odd for even nq) values apart, which makes the
// strided per-thread reads bank-conflict free.
//
// Per output the products are accumulated over p (then q, then r) in ascending
// order starting from 0 with fused multiply-adds -- the same arithmetic, in the
// same order, as every variant of the reference (benchmark04.cc:55-59,
// benchmark05.cc:65-69), so results are bit-identical to the reference kernels.
#pragma once

#include "common.cuh"

namespace b200fe
{

// one row: a[0..NM) (registers) x basis[BOFF + p*NQ + i] -> dst[i*OSTRIDE]
template <typename T, int NM, int NQ, int BOFF, int OSTRIDE>
__device__ __forceinline__ void contract_row_to_smem(const T (&a)[NM], T *__restrict__ dst)
{
#pragma unroll
    for (int i = 0; i < NQ; ++i)
    {
        T t = T(0);
#pragma unroll
        for (int p = 0; p < NM; ++p)
            t = fmadd(a[p], cbasis<T>(BOFF + p * NQ + i), t);
        dst[i * OSTRIDE] = t;
    }
}

template <typename T, int NM, int NQ, int BOFF, int OSTRIDE>
__device__ __forceinline__ void contract_row_to_global(const T (&a)[NM], T *__restrict__ dst)
{
#pragma unroll
    for (int i = 0; i < NQ; ++i)
    {
        T t = T(0);
#pragma unroll
        for (int p = 0; p < NM; ++p)
            t = fmadd(a[p], cbasis<T>(BOFF + p * NQ + i), t);
        st_stream(dst + (size_t)i * OSTRIDE, t);
    }
}

// cooperative contiguous global -> shared copy of `count` values
template <typename T, int THREADS, int MAXCOUNT>
__device__ __forceinline__ void tile_load(T *__restrict__ s, const T *__restrict__ g, int count, bool vec, int tid)
{
    using V         = typename Vec16<T>::type;
    constexpr int W = Vec16<T>::W;
    if (vec)
    {
        const int nv = count / W;
        const V *gv  = reinterpret_cast<const V *>(g);
        V *sv        = reinterpret_cast<V *>(s);
        constexpr int ITER = (MAXCOUNT / W + THREADS - 1) / THREADS;
#pragma unroll
        for (int it = 0; it < ITER; ++it)
        {
            const int c = tid + it * THREADS;
            if (c < nv)
                sv[c] = ld_stream(gv + c);
        }
        for (int c = nv * W + tid; c < count; c += THREADS)
            s[c] = ld_stream(g + c);
    }
    else
    {
        for (int c = tid; c < count; c += THREADS)
            s[c] = ld_stream(g + c);
    }
}

// ============================== quad ==========================================

template <typename T, int NQ, int E, int THREADS> struct QuadRows
{
    static constexpr int NM  = NQ - 1;
    static constexpr int NM2 = NM * NM;
    static constexpr int NQ2 = NQ * NQ;
    static constexpr int W   = Vec16<T>::W;
    // staged output: element stride OS = NQ2 + PAD with OS == NQ (mod one bank
    // row) so that pass 2's lanes (e, i) -- i fastest -- hit distinct banks
    static constexpr int BANKROW = 128 / (int)sizeof(T);
    static constexpr int PAD     = (((NQ - NQ2) % BANKROW) + BANKROW) % BANKROW;
    static constexpr int OS      = NQ2 + PAD;
    // widest vector that divides both the element size and the padded stride
    static constexpr int OVW = (NQ2 % W == 0 && OS % W == 0) ? W : ((NQ2 % 2 == 0 && OS % 2 == 0) ? 2 : 1);
    static constexpr int SA  = E * (OS > NM2 ? OS : NM2); // input tile, later the staged output
    static constexpr int SB  = E * NQ * NM;               // after direction 0: [e][i][q]
    static constexpr size_t SMEM = (size_t)(SA + SB) * sizeof(T);
    static constexpr bool IN_VEC_OK  = ((size_t)E * NM2 * sizeof(T)) % 16 == 0;
    static constexpr bool OUT_VEC_OK = ((size_t)E * NQ2 * sizeof(T)) % 16 == 0;
    static constexpr int B0 = 0, B1 = NM * NQ;
};

template <typename T, int NQ, int E, int THREADS>
__global__ void __launch_bounds__(THREADS)
    bwdtrans_quad_rows_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt, int in_vec, int out_vec)
{
    using C = QuadRows<T, NQ, E, THREADS>;
    constexpr int NM = C::NM, NM2 = C::NM2, NQ2 = C::NQ2, OS = C::OS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    T *sB = sA + C::SA;

    const int tid   = threadIdx.x;
    const size_t e0 = (size_t)blockIdx.x * E;
    const int ne    = (nelmt - e0 < (size_t)E) ? (int)(nelmt - e0) : E;

    tile_load<T, THREADS, E * NM2>(sA, in + e0 * NM2, ne * NM2, in_vec != 0, tid);
    __syncthreads();

    // direction 0: rows (e, q); sA[row*NM + p] -> sB[e][i][q]
    {
        const int nrows    = ne * NM;
        constexpr int ITER = (E * NM + THREADS - 1) / THREADS;
#pragma unroll 1
        for (int it = 0; it < ITER; ++it)
        {
            const int row = tid + it * THREADS;
            if (row < nrows)
            {
                T a[NM];
#pragma unroll
                for (int p = 0; p < NM; ++p)
                    a[p] = sA[row * NM + p];
                const int e = row / NM, q = row - e * NM;
                contract_row_to_smem<T, NM, NQ, C::B0, NM>(a, sB + e * (NQ * NM) + q);
            }
        }
    }
    __syncthreads();

    // direction 1: rows (e, i); sB[row*NM + q] -> staged out sA[e*OS + j*NQ + i]
    {
        const int nrows    = ne * NQ;
        constexpr int ITER = (E * NQ + THREADS - 1) / THREADS;
#pragma unroll 1
        for (int it = 0; it < ITER; ++it)
        {
            const int row = tid + it * THREADS;
            if (row < nrows)
            {
                T a[NM];
#pragma unroll
                for (int q = 0; q < NM; ++q)
                    a[q] = sB[row * NM + q];
                const int e = row / NQ, i = row - e * NQ;
                contract_row_to_smem<T, NM, NQ, C::B1, NQ>(a, sA + e * OS + i);
            }
        }
    }
    __syncthreads();

    // staged tile -> global, contiguous and vectorised
    T *gout = out + e0 * NQ2;
    if (out_vec && C::OVW > 1)
    {
        constexpr int VW  = C::OVW;
        constexpr int CPE = NQ2 / VW; // chunks per element
        const int nchunk  = ne * CPE;
        constexpr int ITER = (E * CPE + THREADS - 1) / THREADS;
#pragma unroll
        for (int it = 0; it < ITER; ++it)
        {
            const int c = tid + it * THREADS;
            if (c < nchunk)
            {
                const int e = c / CPE, m = c - e * CPE;
                if (VW == Vec16<T>::W)
                {
                    using V = typename Vec16<T>::type;
                    st_stream(reinterpret_cast<V *>(gout) + c, *reinterpret_cast<const V *>(sA + e * OS + m * VW));
                }
                else
                {
                    // VW == 2 with T = float
                    st_stream(reinterpret_cast<float2 *>(gout) + c,
                              *reinterpret_cast<const float2 *>(sA + e * OS + m * VW));
                }
            }
        }
    }
    else
    {
        const int n = ne * NQ2;
        for (int c = tid; c < n; c += THREADS)
        {
            const int e = c / NQ2, m = c - e * NQ2;
            st_stream(gout + c, sA[e * OS + m]);
        }
    }
}

// ============================== hex ===========================================

template <typename T, int NQ, int E, int THREADS> struct HexRows
{
    static constexpr int NM  = NQ - 1;
    static constexpr int NM2 = NM * NM;
    static constexpr int NM3 = NM2 * NM;
    static constexpr int NQ2 = NQ * NQ;
    static constexpr int NQ3 = NQ2 * NQ;
    static constexpr int S0  = NM3;      // in            [e][r][q][p]
    static constexpr int S1  = NQ * NM2; // after dir 0   [e][i][r][q]
    static constexpr int S2  = NQ2 * NM; // after dir 1   [e][j][i][r]
    static constexpr int SA  = E * (S2 > S0 ? S2 : S0); // S0, later S2
    static constexpr int SB  = E * S1;
    static constexpr size_t SMEM = (size_t)(SA + SB) * sizeof(T);
    static constexpr bool IN_VEC_OK = ((size_t)E * NM3 * sizeof(T)) % 16 == 0;
    static constexpr int B0 = 0, B1 = NM * NQ, B2 = 2 * NM * NQ;
};

template <typename T, int NQ, int E, int THREADS>
__global__ void __launch_bounds__(THREADS)
    bwdtrans_hex_rows_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt, int in_vec)
{
    using C = HexRows<T, NQ, E, THREADS>;
    constexpr int NM = C::NM, NM2 = C::NM2, NM3 = C::NM3, NQ2 = C::NQ2, NQ3 = C::NQ3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    T *sB = sA + C::SA;

    const int tid   = threadIdx.x;
    const size_t e0 = (size_t)blockIdx.x * E;
    const int ne    = (nelmt - e0 < (size_t)E) ? (int)(nelmt - e0) : E;

    tile_load<T, THREADS, E * NM3>(sA, in + e0 * NM3, ne * NM3, in_vec != 0, tid);
    __syncthreads();

    // direction 0: rows (e, r, q); sA[row*NM + p] -> sB[e][i][r][q]
    {
        const int nrows    = ne * NM2;
        constexpr int ITER = (E * NM2 + THREADS - 1) / THREADS;
#pragma unroll 1
        for (int it = 0; it < ITER; ++it)
        {
            const int row = tid + it * THREADS;
            if (row < nrows)
            {
                T a[NM];
#pragma unroll
                for (int p = 0; p < NM; ++p)
                    a[p] = sA[row * NM + p];
                const int e = row / NM2, rq = row - e * NM2;
                contract_row_to_smem<T, NM, NQ, C::B0, NM2>(a, sB + e * C::S1 + rq);
            }
        }
    }
    __syncthreads();

    // direction 1: rows (e, i, r); sB[row*NM + q] -> sA[e][j][i][r]
    {
        const int nrows    = ne * NQ * NM;
        constexpr int ITER = (E * NQ * NM + THREADS - 1) / THREADS;
#pragma unroll 1
        for (int it = 0; it < ITER; ++it)
        {
            const int row = tid + it * THREADS;
            if (row < nrows)
            {
                T a[NM];
#pragma unroll
                for (int q = 0; q < NM; ++q)
                    a[q] = sB[row * NM + q];
                const int e = row / (NQ * NM), ir = row - e * (NQ * NM);
                contract_row_to_smem<T, NM, NQ, C::B1, NQ * NM>(a, sA + e * C::S2 + ir);
            }
        }
    }
    __syncthreads();

    // direction 2: rows (e, j, i); sA[row*NM + r] -> out[e][k][j][i], a warp
    // writes 32 consecutive values per k (full 128-byte lines)
    {
        const int nrows    = ne * NQ2;
        constexpr int ITER = (E * NQ2 + THREADS - 1) / THREADS;
        T *gout            = out + e0 * NQ3;
#pragma unroll 1
        for (int it = 0; it < ITER; ++it)
        {
            const int row = tid + it * THREADS;
            if (row < nrows)
            {
                T a[NM];
#pragma unroll
                for (int r = 0; r < NM; ++r)
                    a[r] = sA[row * NM + r];
                const int e = row / NQ2, ji = row - e * NQ2;
                contract_row_to_global<T, NM, NQ, C::B2, NQ2>(a, gout + (size_t)e * NQ3 + ji);
            }
        }
    }
}

} // namespace b200fe

// sumfac_iprod.cuh -- IProductWRTBase, the transpose of BwdTrans (named in
// BASELINE.json's north_star, absent from the reference sources):
//   quad  out[e][q][p]    = sum_j B1[q][j] ( sum_i B0[p][i] (w*in)[e][j][i] )
//   hex   out[e][r][q][p] = sum_k B2[r][k] ( sum_j B1[q][j] ( sum_i B0[p][i] (w*in)[e][k][j][i] ) )
// with the quadrature metric w (Jacobian * weights, one value per quadrature point, may be null).
// nq^d values in, nm^d modes out: the same element-batched row passes as the rows back-end of
// BwdTrans (contraction_pass of sumfac_rows.cuh is generic in the row length and the number of
// outputs), with the roles of nm and nq swapped and the basis bank filled TRANSPOSED
// (bank[d][i*nm + p] = B_d[p*nq + i]).  Rows have nq values -- an even count for the swept nq --
// so every shared-memory row is padded to nq + 1 to keep the per-thread strided reads conflict
// free.  Every sum runs in ascending index order from 0 with fused multiply-adds, w*in is one
// rounded product: bit-identical to oracle_iproduct_* (oracle/oracle_impl.h).
#pragma once

#include "sumfac_rows.cuh"

namespace b200fe
{

// cooperative global -> shared copy of `rows` rows of NQ values into rows of stride NQ + 1,
// multiplied by the metric where there is one
template <typename T, int NQ, int THREADS>
__device__ __forceinline__ void iprod_tile_load(T *__restrict__ s, const T *__restrict__ g, const T *__restrict__ w,
                                                int count, int tid)
{
    using V         = typename Vec16<T>::type;
    constexpr int W = Vec16<T>::W;
    // 16-byte loads where the slab allows it (a vector never straddles a row when W divides nq)
    const bool vec = (NQ % W == 0) && ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) &&
                     (!w || (reinterpret_cast<uintptr_t>(w) & 15u) == 0);
    if (vec)
    {
        const int nv = count / W; // count = ne * nq^d is a multiple of W here
        for (int c = tid; c < nv; c += THREADS)
        {
            V v         = ld_stream(reinterpret_cast<const V *>(g) + c);
            T *pv       = reinterpret_cast<T *>(&v);
            if (w)
            {
                const V u   = ld_stream(reinterpret_cast<const V *>(w) + c);
                const T *pu = reinterpret_cast<const T *>(&u);
#pragma unroll
                for (int k = 0; k < W; ++k)
                    pv[k] = pv[k] * pu[k];
            }
            const int e0 = c * W, row = e0 / NQ, col = e0 - row * NQ;
#pragma unroll
            for (int k = 0; k < W; ++k)
                s[row * (NQ + 1) + col + k] = pv[k];
        }
        return;
    }
    for (int c = tid; c < count; c += THREADS)
    {
        const int row = c / NQ, col = c - row * NQ;
        T v           = ld_stream(g + c);
        if (w)
            v = v * ld_stream(w + c);
        s[row * (NQ + 1) + col] = v;
    }
}

// contraction shape: rows fully in registers with immediate constant-bank operands for the small nq,
// the p-loop form beyond (sumfac_rows.cuh, V = 0 / 1)
template <int NQ> constexpr int iprod_v()
{
    return NQ <= 12 ? 0 : 1;
}

template <typename T, int NQ, int E> struct QuadIprodShape
{
    static constexpr int NM = NQ - 1, NQ2 = NQ * NQ, NM2 = NM * NM, RS = NQ + 1;
    static constexpr int S_IN  = E * NQ * RS;  // input tile [e][j][i], rows padded
    static constexpr int S_MID = E * NM * RS;  // after direction 0: [e][p][j], rows padded
    static constexpr int S_OUT = E * NM2;      // staged output [e][q][p] (aliases the input tile)
    static constexpr int SA    = S_IN > S_OUT ? S_IN : S_OUT;
    static constexpr size_t SMEM = (size_t)((SA + 1) / 2 * 2 + S_MID) * sizeof(T);
    static constexpr int B0 = 0, B1 = NQ * bank_pitch<T>(NM); // transposed basis matrices in the bank
};

template <typename T, int NQ, int E, int THREADS, int R>
__global__ void __launch_bounds__(THREADS)
    iproduct_quad_rows_kernel(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out, unsigned nelmt)
{
    using C = QuadIprodShape<T, NQ, E>;
    constexpr int NM = C::NM, RS = C::RS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    T *sB = sA + (C::SA + 1) / 2 * 2;
    const int tid   = threadIdx.x;
    const size_t e0 = (size_t)blockIdx.x * E;
    const int ne    = (nelmt - e0 < (size_t)E) ? (int)(nelmt - e0) : E;

    iprod_tile_load<T, NQ, THREADS>(sA, in + e0 * C::NQ2, w ? w + e0 * C::NQ2 : nullptr, ne * C::NQ2, tid);
    __syncthreads();
    // direction 0: rows (e, j) of nq values -> nm outputs p, to mid[e][p][j]
    contraction_pass<T, NQ, NM, C::B0, RS, THREADS, R, iprod_v<NQ>(), E * NQ, false>(
        ne * NQ, tid, [&](int row) { return sA + row * RS; },
        [&](int row) {
            const int e = row / NQ, j = row - e * NQ;
            return sB + e * (NM * RS) + j;
        });
    __syncthreads();
    // direction 1: rows (e, p) of nq values -> nm outputs q, to the staged out[e][q][p]
    contraction_pass<T, NQ, NM, C::B1, NM, THREADS, R, iprod_v<NQ>(), E * NM, false>(
        ne * NM, tid, [&](int row) { return sB + row * RS; },
        [&](int row) {
            const int e = row / NM, p = row - e * NM;
            return sA + e * C::NM2 + p;
        });
    __syncthreads();
    T *gout     = out + e0 * C::NM2;
    const int n = ne * C::NM2;
    for (int c = tid; c < n; c += THREADS)
        st_stream(gout + c, sA[c]);
}

template <typename T, int NQ, int E> struct HexIprodShape
{
    static constexpr int NM = NQ - 1, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ, NM2 = NM * NM, NM3 = NM2 * NM, RS = NQ + 1;
    static constexpr int S_IN = E * NQ2 * RS;      // input tile [e][k][j][i], rows padded
    static constexpr int S1   = E * NM * NQ * RS;  // after direction 0: [e][p][k][j]
    static constexpr int S2   = E * NM2 * RS;      // after direction 1: [e][q][p][k] (aliases the input tile)
    static constexpr int SA   = S_IN > S2 ? S_IN : S2;
    static constexpr size_t SMEM = (size_t)((SA + 1) / 2 * 2 + S1) * sizeof(T);
    static constexpr int B0 = 0, B1 = NQ * bank_pitch<T>(NM), B2 = 2 * NQ * bank_pitch<T>(NM);
};

template <typename T, int NQ, int E, int THREADS, int R>
__global__ void __launch_bounds__(THREADS)
    iproduct_hex_rows_kernel(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out, unsigned nelmt)
{
    using C = HexIprodShape<T, NQ, E>;
    constexpr int NM = C::NM, RS = C::RS, NQ2 = C::NQ2, NM2 = C::NM2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    T *sB = sA + (C::SA + 1) / 2 * 2;
    const int tid   = threadIdx.x;
    const size_t e0 = (size_t)blockIdx.x * E;
    const int ne    = (nelmt - e0 < (size_t)E) ? (int)(nelmt - e0) : E;

    iprod_tile_load<T, NQ, THREADS>(sA, in + e0 * C::NQ3, w ? w + e0 * C::NQ3 : nullptr, ne * C::NQ3, tid);
    __syncthreads();
    // direction 0: rows (e, k, j) -> outputs p, to s1[e][p][k][j]
    contraction_pass<T, NQ, NM, C::B0, NQ * RS, THREADS, R, iprod_v<NQ>(), E * NQ2, false>(
        ne * NQ2, tid, [&](int row) { return sA + row * RS; },
        [&](int row) {
            const int e = row / NQ2, kj = row - e * NQ2, k = kj / NQ, j = kj - k * NQ;
            return sB + e * (NM * NQ * RS) + k * RS + j;
        });
    __syncthreads();
    // direction 1: rows (e, p, k) -> outputs q, to s2[e][q][p][k]
    contraction_pass<T, NQ, NM, C::B1, NM * RS, THREADS, R, iprod_v<NQ>(), E * NM * NQ, false>(
        ne * NM * NQ, tid, [&](int row) { return sB + row * RS; },
        [&](int row) {
            const int e = row / (NM * NQ), pk = row - e * (NM * NQ), p = pk / NQ, k = pk - p * NQ;
            return sA + e * (NM2 * RS) + p * RS + k;
        });
    __syncthreads();
    // direction 2: rows (e, q, p) -> outputs r, straight to out[e][r][q][p] (lanes along p: coalesced)
    T *gout = out + e0 * C::NM3;
    contraction_pass<T, NQ, NM, C::B2, NM2, THREADS, R, iprod_v<NQ>(), E * NM2, true>(
        ne * NM2, tid, [&](int row) { return sA + row * RS; },
        [&](int row) {
            const int e = row / NM2, qp = row - e * NM2;
            return gout + (size_t)e * C::NM3 + qp;
        });
}

} // namespace b200fe

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

// ---- persistent, TMA-fed twin of the hex row kernel ("iprod-pipe") ---------------------------------------------------
// The row kernel above fetches its tile with plain loads and scatters it into padded rows: one tile per CTA, the load
// fully exposed (hex nq = 10 FP64: 0.58 of the roofline).  Here the CTAs are persistent and the contiguous slab of the
// tile after next (and of the metric) is fetched by a 1-D bulk copy into a two-slot ring while the current tile is
// contracted, exactly like bwdtrans_hex_pipe_kernel (sumfac_rows.cuh).  The slab stays unpadded -- rows of nq values,
// a 2-way conflicted read pattern for even nq, paid once per tile in direction 0 only -- and the metric is multiplied
// in place in the slot by a vectorised sweep before direction 0 (one rounded product, as in the row kernel).
template <typename T, int NQ, int E, bool WEIGHTED> struct HexIprodPipe
{
    static constexpr int NM = NQ - 1, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ, NM2 = NM * NM, NM3 = NM2 * NM, RS = NQ + 1;
    static constexpr int SLOT = (E * NQ3 * (int)sizeof(T) + 15) / 16 * 16 / (int)sizeof(T);
    static constexpr int S1   = (E * NM * NQ * RS + 1) / 2 * 2; // after direction 0: [e][p][k][j], rows padded
    static constexpr int S2   = (E * NM2 * RS + 1) / 2 * 2;     // after direction 1: [e][q][p][k]
    static constexpr int NSL  = WEIGHTED ? 4 : 2;
    static constexpr size_t SMEM = 32 + (size_t)(NSL * SLOT + S1 + S2) * sizeof(T);
    static constexpr int B0 = 0, B1 = NQ * bank_pitch<T>(NM), B2 = 2 * NQ * bank_pitch<T>(NM);
    static_assert(((size_t)E * NQ3 * sizeof(T)) % 16 == 0, "pipe tiles must be 16-byte granular");
};

template <typename T, int NQ, int E, int THREADS, int R, bool WEIGHTED>
__device__ __noinline__ void iproduct_hex_pipe_body(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out,
                                                    unsigned nelmt, unsigned ntiles);

template <typename T, int NQ, int E, int THREADS, int R, bool WEIGHTED>
__global__ void __launch_bounds__(THREADS)
    iproduct_hex_pipe_kernel(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out, unsigned nelmt,
                             unsigned ntiles)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    iproduct_hex_pipe_body<T, NQ, E, THREADS, R, WEIGHTED>(in, w, out, nelmt, ntiles);
}

template <typename T, int NQ, int E, int THREADS, int R, bool WEIGHTED>
__device__ __noinline__ void iproduct_hex_pipe_body(const T *__restrict__ in, const T *__restrict__ w, T *__restrict__ out,
                                                    unsigned nelmt, unsigned ntiles)
{
    using C = HexIprodPipe<T, NQ, E, WEIGHTED>;
    constexpr int NM = C::NM, RS = C::RS, NQ2 = C::NQ2, NM2 = C::NM2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw); // [0..1]: in slots, [2..3]: metric slots
    T *slot0      = reinterpret_cast<T *>(smem_raw + 32);
    T *wslot0     = slot0 + 2 * C::SLOT;
    T *s1         = slot0 + C::NSL * C::SLOT;
    T *s2         = s1 + C::S1;
    const int tid = threadIdx.x;

    if (tid == 0)
    {
#pragma unroll
        for (int b = 0; b < 4; ++b)
            mbar_init(&bar[b], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](unsigned t, unsigned b) {
        ring_issue<T, E, C::NQ3>(slot0 + b * C::SLOT, &bar[b], in, t, nelmt);
        if (WEIGHTED)
            ring_issue<T, E, C::NQ3>(wslot0 + b * C::SLOT, &bar[2 + b], w, t, nelmt);
    };
    if (tid == 0)
        for (unsigned s = 0; s < 2; ++s)
        {
            const unsigned t = blockIdx.x + s * gridDim.x;
            if (t < ntiles)
                issue(t, s);
        }

    unsigned it = 0;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it)
    {
        const unsigned b = it & 1u;
        T *s_in          = slot0 + b * C::SLOT;
        const size_t e0  = (size_t)tile * E;
        const int ne     = (nelmt - e0 < (size_t)E) ? (int)(nelmt - e0) : E;
        ring_wait<T, C::NQ3>(s_in, &bar[b], (it >> 1) & 1u, in + e0 * C::NQ3, ne, tid);
        if constexpr (WEIGHTED)
        {
            T *s_w = wslot0 + b * C::SLOT;
            ring_wait<T, C::NQ3>(s_w, &bar[2 + b], (it >> 1) & 1u, w + e0 * C::NQ3, ne, tid);
            using V         = typename Vec16<T>::type;
            constexpr int W = Vec16<T>::W;
            const int nv    = (ne * C::NQ3 + W - 1) / W; // the slot is padded to whole vectors
            for (int c = tid; c < nv; c += THREADS)
            {
                V v       = reinterpret_cast<V *>(s_in)[c];
                const V u = reinterpret_cast<const V *>(s_w)[c];
                T *pv     = reinterpret_cast<T *>(&v);
                const T *pu = reinterpret_cast<const T *>(&u);
#pragma unroll
                for (int k = 0; k < W; ++k)
                    pv[k] = pv[k] * pu[k];
                reinterpret_cast<V *>(s_in)[c] = v;
            }
            __syncthreads();
        }
        // direction 0: rows (e, k, j) -> outputs p, to s1[e][p][k][j]
        contraction_pass<T, NQ, NM, C::B0, NQ * RS, THREADS, R, iprod_v<NQ>(), E * NQ2, false>(
            ne * NQ2, tid, [&](int row) { return s_in + row * NQ; },
            [&](int row) {
                const int e = row / NQ2, kj = row - e * NQ2, k = kj / NQ, j = kj - k * NQ;
                return s1 + e * (NM * NQ * RS) + k * RS + j;
            });
        __syncthreads(); // the slots are drained: refill them with the tile after next
        if (tid == 0)
        {
            const unsigned nxt = tile + 2 * gridDim.x;
            if (nxt < ntiles)
                issue(nxt, b);
        }
        // direction 1: rows (e, p, k) -> outputs q, to s2[e][q][p][k]
        contraction_pass<T, NQ, NM, C::B1, NM * RS, THREADS, R, iprod_v<NQ>(), E * NM * NQ, false>(
            ne * NM * NQ, tid, [&](int row) { return s1 + row * RS; },
            [&](int row) {
                const int e = row / (NM * NQ), pk = row - e * (NM * NQ), p = pk / NQ, k = pk - p * NQ;
                return s2 + e * (NM2 * RS) + p * RS + k;
            });
        __syncthreads();
        // direction 2: rows (e, q, p) -> outputs r, straight to out[e][r][q][p] (lanes along p: coalesced)
        T *gout = out + e0 * C::NM3;
        contraction_pass<T, NQ, NM, C::B2, NM2, THREADS, R, iprod_v<NQ>(), E * NM2, true>(
            ne * NM2, tid, [&](int row) { return s2 + row * RS; },
            [&](int row) {
                const int e = row / NM2, qp = row - e * NM2;
                return gout + (size_t)e * C::NM3 + qp;
            });
        // s1 is rewritten by the next direction-0 pass (every warp has passed the barrier after direction 1); s2 by the
        // next direction-1 pass, entered only after the barrier that follows the next direction-0 pass.
    }
}

} // namespace b200fe

// sumfac_rows_coa.cuh -- the rows back-end on the warp-interleaved layout of the reference's
// "Coales" kernels (benchmark04.cc:78-147, benchmark05.cc:104-201):
//     x[(e/32)*32*len + 32*idx + e%32]
// for the nq the thread-per-element back-end (sumfac_tpe.cuh) cannot hold in registers.
//
// A CTA takes E consecutive elements of one group of 32 (E divides 32, E*sizeof(T) >= 32 bytes,
// so every global access is a whole 32-byte sector): the gather turns the interleaved slab into
// the element-major shared-memory tile the rows passes expect, the passes are the ones of
// sumfac_rows.cuh, and the last step scatters back with the lanes running over the elements
// (consecutive addresses within a sector).  The other sectors of each 128-byte line are read /
// written by the neighbouring CTAs of the same group at about the same time, so the lines meet
// in L2 and HBM traffic stays at the algorithmic bytes.
#pragma once

#include "sumfac_rows.cuh"

namespace b200fe
{

// interleaved global -> element-major shared tile: s[el*LEN + idx] = g[32*idx + el], el < E.
// All of a thread's loads are issued before its first store (the loop is latency-bound otherwise:
// profiles/r01_ncu_hex8_f64_coa.txt, 44 % of the stall samples on the store waiting for its load), and
// they are 16-byte loads (W elements of the same idx) where the slab is 16-byte aligned.
template <typename T, int E, int LEN, int THREADS>
__device__ __forceinline__ void coa_gather(T *__restrict__ s, const T *__restrict__ g, int tid)
{
    using V         = typename Vec16<T>::type;
    constexpr int W = Vec16<T>::W;
    static_assert(E % W == 0, "a tile holds whole 16-byte vectors of elements");
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0)
    {
        constexpr int PARTS = E / W, N = PARTS * LEN, ITER = (N + THREADS - 1) / THREADS;
        V v[ITER];
#pragma unroll
        for (int it = 0; it < ITER; ++it)
        {
            const int c = tid + it * THREADS;
            if (c < N)
            {
                const int idx = c / PARTS, part = c - idx * PARTS;
                v[it]         = ld_stream(reinterpret_cast<const V *>(g + 32 * idx + part * W));
            }
        }
#pragma unroll
        for (int it = 0; it < ITER; ++it)
        {
            const int c = tid + it * THREADS;
            if (c < N)
            {
                const int idx = c / PARTS, part = c - idx * PARTS;
                const T *pv   = reinterpret_cast<const T *>(&v[it]);
#pragma unroll
                for (int k = 0; k < W; ++k)
                    s[(part * W + k) * LEN + idx] = pv[k];
            }
        }
        return;
    }
    constexpr int N = E * LEN;
    for (int c = tid; c < N; c += THREADS)
    {
        const int idx = c / E, el = c - idx * E;
        s[el * LEN + idx] = ld_stream(g + 32 * idx + el);
    }
}

// quad: the element-major rows kernel with the gather in front and a scatter of the staged tile behind
template <typename T, int NQ, int E, int THREADS, int R, int V>
__device__ __noinline__ void bwdtrans_quad_rowscoa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int E, int THREADS, int R, int V>
__global__ void __launch_bounds__(THREADS)
    bwdtrans_quad_rowscoa_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    bwdtrans_quad_rowscoa_body<T, NQ, E, THREADS, R, V>(in, out, nelmt);
}

template <typename T, int NQ, int E, int THREADS, int R, int V>
__device__ __noinline__ void bwdtrans_quad_rowscoa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    using C = QuadRows<T, NQ, E, THREADS, R, V>;
    static_assert(32 % E == 0, "a tile must not straddle interleave groups");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    T *sB = sA + C::align16(C::SA);
    const int tid        = threadIdx.x;
    constexpr int PER    = 32 / E;                      // tiles per group
    const unsigned group = blockIdx.x / PER, l0 = (blockIdx.x % PER) * E;
    const T *gin         = in + (size_t)group * 32 * C::NM2 + l0;
    T *gout              = out + (size_t)group * 32 * C::NQ2 + l0;
    (void)nelmt; // nelmt % 32 == 0 is checked at the C ABI: every tile is full

    coa_gather<T, E, C::NM2, THREADS>(sA, gin, tid);
    __syncthreads();
    quad_dir0<T, NQ, E, THREADS, R, V>(sA, sB, E, tid);
    __syncthreads();
    quad_dir1<T, NQ, E, THREADS, R, V>(sB, sA, E, tid);
    __syncthreads();
    constexpr int N = E * C::NQ2;
    for (int c = tid; c < N; c += THREADS)
    {
        const int m = c / E, el = c - m * E;
        st_stream(gout + 32 * m + el, sA[el * C::OS + m]);
    }
}

// hex: directions 0 and 1 as in the element-major kernel; direction 2 with the rows re-indexed so that
// consecutive lanes hold consecutive elements of the same (j, i), and stored interleaved
template <typename T, int NQ, int E, int THREADS, int R, int V>
__device__ __noinline__ void bwdtrans_hex_rowscoa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int E, int THREADS, int R, int V>
__global__ void __launch_bounds__(THREADS)
    bwdtrans_hex_rowscoa_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    bwdtrans_hex_rowscoa_body<T, NQ, E, THREADS, R, V>(in, out, nelmt);
}

template <typename T, int NQ, int E, int THREADS, int R, int V>
__device__ __noinline__ void bwdtrans_hex_rowscoa_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    using C = HexRows<T, NQ, E, THREADS, R, V>;
    static_assert(32 % E == 0, "a tile must not straddle interleave groups");
    constexpr int NM = C::NM, NQ2 = C::NQ2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sA = reinterpret_cast<T *>(smem_raw);
    T *sB = sA + C::align16(C::SA);
    const int tid        = threadIdx.x;
    constexpr int PER    = 32 / E;
    const unsigned group = blockIdx.x / PER, l0 = (blockIdx.x % PER) * E;
    const T *gin         = in + (size_t)group * 32 * C::NM3 + l0;
    T *gout              = out + (size_t)group * 32 * C::NQ3 + l0;
    (void)nelmt;

    coa_gather<T, E, C::NM3, THREADS>(sA, gin, tid);
    __syncthreads();
    hex_dir0<T, NQ, E, THREADS, R, V>(sA, sB, E, tid);
    __syncthreads();
    hex_dir1<T, NQ, E, THREADS, R, V>(sB, sA, E, tid);
    __syncthreads();
    // rows (j, i, e) with e fastest: s2[e][j][i][r] -> out_coa[k][j][i][e]
    contraction_pass<T, NM, NQ, C::B2, 32 * NQ2, THREADS, R, V, E * NQ2, true>(
        E * NQ2, tid,
        [&](int row) {
            const int ji = row / E, e = row - ji * E;
            return sA + e * C::ES2 + ji * NM;
        },
        [&](int row) {
            const int ji = row / E, e = row - ji * E;
            return gout + 32 * ji + e;
        });
}

} // namespace b200fe

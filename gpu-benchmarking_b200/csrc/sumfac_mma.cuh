// sumfac_mma.cuh -- "mma" back-end: FP64 sum-factorisation on the tensor cores
// (DMMA, mma.sync.m8n8k4.f64) for the large-nq quads, where the contraction is a
// dense GEMM-shaped problem that the scalar FP64 pipe cannot feed fast enough:
// per element 2*nq*nm*(nm+nq) flops against 8*(nm^2+nq^2) bytes is 3.4-7.9 flop/B
// for nq = 14..32, i.e. 22-51 TFLOP/s at the HBM roofline, while DFMA with a
// constant-bank operand tops out near 27 TFLOP/s on B200 (profiles/r01_ubench_fp_pipe.txt)
// and DMMA sustains 37 TFLOP/s from registers with 1/8 of the issue slots
// (profiles/r01_ubench_dmma.txt).
//
// Work unit: one WARP owns a group of G consecutive elements from load to store;
// warps never synchronise with each other (only __syncwarp), so their load /
// contract / store phases drift apart and overlap on the SM.
//
//   load   the group's contiguous slab in[e0 .. e0+G) is fetched by ONE bulk
//          tensor-memory-accelerator copy (cp.async.bulk, completion on the warp's
//          own mbarrier).  The copy of the next group is issued as soon as direction
//          0 has consumed the slot, so it lands while direction 1 computes.
//   dir 0  mid[(e,q)][i] = sum_p in[(e,q)][p] * B0[p][i]
//          A = data rows (M = flattened (e,q), K = p), B = basis fragments.
//   dir 1  out[e][j][i]  = sum_q B1[q][j] * mid[e][q][i]
//          A = transposed basis fragments (M = j, K = q), B = mid (N = flattened
//          (e,i)); each lane ends up with out[e][j][i..i+1]: one 16-byte store.
//
// Padding of M/N/K up to the 8x8x4 tile is done with zeros in the basis
// fragments and with clamped / zero-selected data loads, never by touching
// memory outside the group.  The summation order inside a k = 4 step is the
// hardware's, so results agree with the reference to rounding (<= 1e-12
// relative, tests/test_bwdtrans_gpu.py), not bit for bit; the rows / pipe
// back-ends remain the bit-exact ones.
#pragma once

#include "sumfac_rows.cuh"

namespace b200fe
{

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// row stride of the intermediate: even (16-byte row starts) and == 4 or 12 (mod 16)
// so that the four k-rows a direction-1 B fragment touches fall into distinct banks
constexpr int mma_mid_stride(int nq)
{
    int s = nq + (nq & 1);
    while (s % 16 != 4 && s % 16 != 12)
        s += 2;
    return s;
}

template <int NQ, int G, int WARPS, int MB0, int NB1> struct QuadMma
{
    static constexpr int NM    = NQ - 1;
    static constexpr int NM2   = NM * NM;
    static constexpr int NQ2   = NQ * NQ;
    static constexpr int KS    = (NM + 3) / 4;     // k steps (K = nm in both directions)
    static constexpr int NT0   = (NQ + 7) / 8;     // direction 0: n tiles over i
    static constexpr int MT1   = (NQ + 7) / 8;     // direction 1: m tiles over j
    static constexpr int ROWS0 = G * NM;           // direction 0: M = flattened (e, q)
    static constexpr int MT0   = (ROWS0 + 7) / 8;
    static constexpr int COLS1 = G * NQ;           // direction 1: N = flattened (e, i)
    static constexpr int NT1   = (COLS1 + 7) / 8;
    static constexpr int S     = mma_mid_stride(NQ);
    static constexpr int SLOT  = (G * NM2 + 1 + 3 + 1) / 2 * 2; // +1: 8-byte window offset, +3: k over-read of the last row
    static constexpr int MID   = G * NM * S;
    static constexpr int WARP_D = SLOT + MID;      // doubles per warp (even)
    static constexpr int FRAG0 = KS * NT0 * 32;
    static constexpr int FRAG1 = MT1 * KS * 32;
    static constexpr int BAR_BYTES = (WARPS * 8 + 15) / 16 * 16;
    static constexpr size_t SMEM = BAR_BYTES + (size_t)(FRAG0 + FRAG1 + WARPS * WARP_D) * sizeof(double);
    static_assert(NQ % 2 == 0, "the mma back-end pairs outputs along i");
};

// Fetch group gn into the warp's slot.  Returns true when the data arrives through the
// mbarrier (bulk copy of the enclosing 16-byte aligned window), false when it was copied
// with ordinary loads (window would leave the array: first group of a misaligned array,
// last group of an array whose end is not 16-byte aligned).
template <int G, int NM2>
__device__ __forceinline__ bool mma_fetch_group(double *slot, uint64_t *bar, const double *__restrict__ in,
                                                unsigned gn, unsigned nelmt, int lane)
{
    const size_t e0      = (size_t)gn * G;
    const unsigned ne    = (nelmt - e0 < (size_t)G) ? (unsigned)(nelmt - e0) : (unsigned)G;
    const unsigned count = ne * (unsigned)NM2;
    const double *src    = in + e0 * NM2;
    const unsigned off   = (unsigned)((reinterpret_cast<uintptr_t>(src) & 15u) >> 3);
    const double *wsrc   = src - off;
    const unsigned bytes = ((off + count) * 8u + 15u) & ~15u;
    const bool fits      = (wsrc >= in) && (reinterpret_cast<const char *>(wsrc) + bytes <=
                                       reinterpret_cast<const char *>(in + (size_t)nelmt * NM2));
    if (fits)
    {
        if (lane == 0)
        {
            fence_proxy_async(); // the slot's earlier generic-proxy reads are ordered before the async write
            mbar_arrive_expect_tx(bar, bytes);
            bulk_load(slot, wsrc, bytes, bar);
        }
        return true;
    }
    for (unsigned k = lane; k < count; k += 32)
        slot[off + k] = src[k];
    __syncwarp();
    return false;
}

template <int NQ, int G, int WARPS, int MB0, int NB1>
__global__ void __launch_bounds__(WARPS * 32)
    bwdtrans_quad_mma_kernel(const double *__restrict__ basis0, const double *__restrict__ basis1,
                             const double *__restrict__ in, double *__restrict__ out, unsigned nelmt, unsigned ngroups,
                             int out_vec)
{
    using C = QuadMma<NQ, G, WARPS, MB0, NB1>;
    constexpr int NM = C::NM, KS = C::KS, NT0 = C::NT0, MT1 = C::MT1, S = C::S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    double *frag0  = reinterpret_cast<double *>(smem_raw + C::BAR_BYTES);
    double *frag1  = frag0 + C::FRAG0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane >> 2, c = lane & 3;
    double *slot = frag1 + C::FRAG1 + warp * C::WARP_D;
    double *mid  = slot + C::SLOT;

    // basis matrices in fragment order, zero padded to the tile grid
    for (int idx = threadIdx.x; idx < C::FRAG0; idx += WARPS * 32)
    {
        const int l = idx & 31, t = idx >> 5, nt = t % NT0, ks = t / NT0;
        const int p = 4 * ks + (l & 3), i = 8 * nt + (l >> 2);
        frag0[idx]  = (p < NM && i < NQ) ? basis0[p * NQ + i] : 0.0;
    }
    for (int idx = threadIdx.x; idx < C::FRAG1; idx += WARPS * 32)
    {
        const int l = idx & 31, t = idx >> 5, ks = t % KS, mt = t / KS;
        const int j = 8 * mt + (l >> 2), q = 4 * ks + (l & 3);
        frag1[idx]  = (j < NQ && q < NM) ? basis1[q * NQ + j] : 0.0;
    }
    if (lane == 0)
    {
        mbar_init(&bars[warp], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned nw = gridDim.x * WARPS;
    unsigned g        = blockIdx.x * WARPS + warp;
    unsigned phase    = 0;
    bool by_barrier   = false;
    if (g < ngroups)
        by_barrier = mma_fetch_group<G, C::NM2>(slot, &bars[warp], in, g, nelmt, lane);

    for (; g < ngroups; g += nw)
    {
        const size_t e0 = (size_t)g * G;
        const int ne    = (nelmt - e0 < (size_t)G) ? (int)(nelmt - e0) : G;
        const double *s_in =
            slot + ((reinterpret_cast<uintptr_t>(in + e0 * C::NM2) & 15u) >> 3); // window offset of this group
        if (by_barrier)
        {
            mbar_wait(&bars[warp], phase);
            phase ^= 1u;
        }

        // ---- direction 0: mid[(e,q)][i] ----------------------------------------------------
        const int nrows = ne * NM;
#pragma unroll 1
        for (int mt = 0; mt < C::MT0; mt += MB0)
        {
            if (mt * 8 >= nrows)
                break;
            double acc[MB0][NT0][2];
            const double *ap[MB0];
#pragma unroll
            for (int m = 0; m < MB0; ++m)
            {
                const int row = (mt + m) * 8 + r;
                ap[m]         = s_in + (row < nrows ? row : nrows - 1) * NM + c;
#pragma unroll
                for (int n = 0; n < NT0; ++n)
                    acc[m][n][0] = acc[m][n][1] = 0.0;
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
            {
                double a[MB0], b[NT0];
#pragma unroll
                for (int m = 0; m < MB0; ++m)
                {
                    a[m] = ap[m][4 * ks]; // over-reads at most 3 values past the row: still inside the slot
                    if (4 * ks + 3 >= NM && 4 * ks + c >= NM)
                        a[m] = 0.0;
                }
#pragma unroll
                for (int n = 0; n < NT0; ++n)
                    b[n] = frag0[(ks * NT0 + n) * 32 + lane];
#pragma unroll
                for (int m = 0; m < MB0; ++m)
#pragma unroll
                    for (int n = 0; n < NT0; ++n)
                        dmma884(acc[m][n], a[m], b[n]);
            }
#pragma unroll
            for (int m = 0; m < MB0; ++m)
            {
                const int row = (mt + m) * 8 + r;
                if (row < nrows)
                {
                    double *mp = mid + row * S + 2 * c;
#pragma unroll
                    for (int n = 0; n < NT0; ++n)
                        if (8 * n + 2 * c < NQ)
                            *reinterpret_cast<double2 *>(mp + 8 * n) = make_double2(acc[m][n][0], acc[m][n][1]);
                }
            }
        }
        __syncwarp();

        // the slot is drained: start fetching this warp's next group under direction 1
        if (g + nw < ngroups)
            by_barrier = mma_fetch_group<G, C::NM2>(slot, &bars[warp], in, g + nw, nelmt, lane);

        // ---- direction 1: out[e][j][i] -----------------------------------------------------
        const int ncols = ne * NQ;
        double *gout    = out + e0 * C::NQ2;
#pragma unroll 1
        for (int nt = 0; nt < C::NT1; nt += NB1)
        {
            if (nt * 8 >= ncols)
                break;
            double acc[MT1][NB1][2];
            const double *bp[NB1];
#pragma unroll
            for (int t = 0; t < NB1; ++t)
            {
                const int n  = (nt + t) * 8 + r;
                const int nc = n < ncols ? n : ncols - 1;
                const int e = nc / NQ, i = nc - e * NQ;
                bp[t] = mid + (e * NM + c) * S + i;
#pragma unroll
                for (int m = 0; m < MT1; ++m)
                    acc[m][t][0] = acc[m][t][1] = 0.0;
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
            {
                double a[MT1], b[NB1];
#pragma unroll
                for (int m = 0; m < MT1; ++m)
                    a[m] = frag1[(m * KS + ks) * 32 + lane];
#pragma unroll
                for (int t = 0; t < NB1; ++t)
                {
                    if (4 * ks + 3 < NM)
                        b[t] = bp[t][4 * ks * S];
                    else
                        b[t] = (4 * ks + c < NM) ? bp[t][4 * ks * S] : 0.0; // predicated: rows >= nm do not exist
                }
#pragma unroll
                for (int m = 0; m < MT1; ++m)
#pragma unroll
                    for (int t = 0; t < NB1; ++t)
                        dmma884(acc[m][t], a[m], b[t]);
            }
#pragma unroll
            for (int t = 0; t < NB1; ++t)
            {
                const int n = (nt + t) * 8 + 2 * c;
                if (n < ncols)
                {
                    const int e = n / NQ, i = n - e * NQ;
                    double *op = gout + (size_t)e * C::NQ2 + i;
#pragma unroll
                    for (int m = 0; m < MT1; ++m)
                    {
                        const int j = 8 * m + r;
                        if (j < NQ)
                        {
                            if (out_vec)
                                st_stream(reinterpret_cast<double2 *>(op + j * NQ),
                                          make_double2(acc[m][t][0], acc[m][t][1]));
                            else
                            {
                                st_stream(op + j * NQ, acc[m][t][0]);
                                st_stream(op + j * NQ + 1, acc[m][t][1]);
                            }
                        }
                    }
                }
            }
        }
        __syncwarp(); // mid is rewritten by the next group's direction 0
    }
}

} // namespace b200fe

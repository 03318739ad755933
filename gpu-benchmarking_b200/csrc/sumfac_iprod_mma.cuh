// sumfac_iprod_mma.cuh -- IProductWRTBase on the FP64 tensor cores (DMMA), the transpose twin of
// sumfac_mma.cuh: nq^d quadrature values in, nm^d modes out.
//
//   dir 0   s1[(e,..,j)][p] = sum_i (w*in)[(e,..,j)][i] B0[p][i]      A = data rows (K = nq: no k padding for
//                                                                     nq % 4 == 0), B = basis fragments (N = p)
//   dir 1+  s[..][q][..]    = sum_j B1[q][j] s1[..][j][..]            A = basis fragments (M = q), B = intermediate
//   last    the result is staged in shared memory with p (and q) padded to the tile width and copied out
//           unpadded, contiguous (nm^d values per element, odd: no 16-byte pairing possible)
//
// Same work split as the BwdTrans kernels: one warp owns a group of G elements from load to store, the group's
// slab (and the metric's, when there is one) arrives by bulk (TMA) copies on the warp's mbarrier, the next
// group's copies are issued as soon as direction 0 has drained the slots.  Sums run over ascending indices
// with fused multiply-adds (DMMA.8x8x4 accumulates in k order): bit-identical to oracle_iproduct_*.
#pragma once

#include "sumfac_mma.cuh"

namespace b200fe
{

constexpr int ipm_pad8(int n)
{
    return (n + 7) / 8 * 8;
}
// stride >= n, even, == 4 or 12 (mod 16): the four k rows of a B fragment fall into distinct banks
constexpr int ipm_stride(int n)
{
    int s = n + (n & 1);
    while (s % 16 != 4 && s % 16 != 12)
        s += 2;
    return s;
}

// rows of KLEN values at src + row*KLEN (raw slab), optionally times the metric at wsrc + row*KLEN;
// output (row, n) for n < NOUTP (padded output count, zero columns beyond NOUT) to dst + row*DS + n
template <int KLEN, int NOUT, int DS, int MB, int NROWS, bool WEIGHTED>
__device__ __forceinline__ void ipm_pass_data(const double *__restrict__ src, const double *__restrict__ wsrc,
                                              const double *__restrict__ fragB, double *__restrict__ dst, int nrows,
                                              int lane)
{
    constexpr int KS = (KLEN + 3) / 4, NT = (NOUT + 7) / 8, MT = (NROWS + 7) / 8, NBLK = (MT + MB - 1) / MB;
    const int r = lane >> 2, c = lane & 3;
    const double *abase = src + r * KLEN + c;
    const double *wbase = wsrc + r * KLEN + c;
    double *dbase       = dst + r * DS + 2 * c;
    const bool kpad     = c >= KLEN - 4 * (KS - 1);
    double b[KS][NT];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int n = 0; n < NT; ++n)
            b[ks][n] = fragB[(ks * NT + n) * 32 + lane];
    double a[2][MB][KS];
    auto load_block = [&](int blk, double (&dstA)[MB][KS]) {
#pragma unroll
        for (int m = 0; m < MB; ++m)
            if (blk * MB + m < MT)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                {
                    const int off = (blk * MB + m) * 8 * KLEN + 4 * ks;
                    double v      = abase[off]; // rows past the group's last read stale data of this warp's slot
                    if (WEIGHTED)
                        v = v * wbase[off];
                    if (4 * ks + 3 >= KLEN && kpad)
                        v = 0.0;
                    dstA[m][ks] = v;
                }
    };
    load_block(0, a[0]);
#pragma unroll
    for (int blk = 0; blk < NBLK; ++blk)
    {
        if (blk + 1 < NBLK)
            load_block(blk + 1, a[(blk + 1) & 1]);
        double acc[MB][NT][2];
#pragma unroll
        for (int m = 0; m < MB; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n)
                acc[m][n][0] = acc[m][n][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int m = 0; m < MB; ++m)
                if (blk * MB + m < MT)
#pragma unroll
                    for (int n = 0; n < NT; ++n)
                        dmma884_ordered(acc[m][n], a[blk & 1][m][ks], b[ks][n]);
#pragma unroll
        for (int m = 0; m < MB; ++m)
            if (blk * MB + m < MT)
            {
                const int row = (blk * MB + m) * 8 + r;
                if (row < nrows)
#pragma unroll
                    for (int n = 0; n < NT; ++n)
                        *reinterpret_cast<double2 *>(dbase + (blk * MB + m) * 8 * DS + 8 * n) =
                            make_double2(acc[m][n][0], acc[m][n][1]);
            }
    }
}

// A = basis fragments (M = MOUT outputs, K = KLEN); column n = (g, w), w < W (W % 8 == 0), its K values at
// src + (g*KLEN + k)*STRIDE + w; output (m, n) to dst + g*DG + m*DM + w for m < 8*MT (padded rows are zeros)
template <int MOUT, int KLEN, int W, int STRIDE, int DG, int DM, int NB, int NCOLS>
__device__ __forceinline__ void ipm_pass_basis(const double *__restrict__ src, const double *__restrict__ fragA,
                                               double *__restrict__ dst, int ncols, int lane)
{
    constexpr int KS = (KLEN + 3) / 4, MT = (MOUT + 7) / 8, NTT = NCOLS / 8, NBLK = (NTT + NB - 1) / NB;
    static_assert(W % 8 == 0 && NCOLS % 8 == 0, "whole column tiles per group");
    const int r = lane >> 2, c = lane & 3;
    const double *bbase = src + c * STRIDE + r;
    double *dbase       = dst + r * DM + 2 * c;
    const bool kpad     = c >= KLEN - 4 * (KS - 1);
    double a[MT][KS];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
            a[m][ks] = fragA[(m * KS + ks) * 32 + lane];
    double b[2][NB][KS];
    auto load_block = [&](int blk, double (&dstB)[NB][KS]) {
#pragma unroll
        for (int t = 0; t < NB; ++t)
            if (blk * NB + t < NTT)
            {
                const int g = ((blk * NB + t) * 8) / W, w0 = ((blk * NB + t) * 8) % W;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                {
                    if (4 * ks + 3 < KLEN)
                        dstB[t][ks] = bbase[(g * KLEN + 4 * ks) * STRIDE + w0];
                    else
                        dstB[t][ks] = kpad ? 0.0 : bbase[(g * KLEN + 4 * ks) * STRIDE + w0];
                }
            }
    };
    load_block(0, b[0]);
#pragma unroll
    for (int blk = 0; blk < NBLK; ++blk)
    {
        if (blk + 1 < NBLK)
            load_block(blk + 1, b[(blk + 1) & 1]);
        double acc[MT][NB][2];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int t = 0; t < NB; ++t)
                acc[m][t][0] = acc[m][t][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int t = 0; t < NB; ++t)
                    if (blk * NB + t < NTT)
                        dmma884_ordered(acc[m][t], a[m][ks], b[blk & 1][t][ks]);
#pragma unroll
        for (int t = 0; t < NB; ++t)
            if (blk * NB + t < NTT)
            {
                const int n0 = (blk * NB + t) * 8, g = n0 / W, w0 = n0 % W;
                if (n0 + 2 * c < ncols)
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        *reinterpret_cast<double2 *>(dbase + g * DG + w0 + 8 * m * DM) =
                            make_double2(acc[m][t][0], acc[m][t][1]);
            }
    }
}

// basis B[p*NQ + i] -> B-operand fragments of its transpose [ks][nt][lane]: (k = i, n = p)
template <int NQ, int THREADS>
__device__ __forceinline__ void ipm_fill_fragB(double *__restrict__ frag, const double *__restrict__ basis)
{
    constexpr int NM = NQ - 1, KS = (NQ + 3) / 4, NT = (NM + 7) / 8;
    for (int idx = threadIdx.x; idx < KS * NT * 32; idx += THREADS)
    {
        const int l = idx & 31, t = idx >> 5, nt = t % NT, ks = t / NT;
        const int i = 4 * ks + (l & 3), p = 8 * nt + (l >> 2);
        frag[idx]   = (p < NM && i < NQ) ? basis[p * NQ + i] : 0.0;
    }
}
// basis B[q*NQ + j] -> A-operand fragments [mt][ks][lane]: (m = q, k = j)
template <int NQ, int THREADS>
__device__ __forceinline__ void ipm_fill_fragA(double *__restrict__ frag, const double *__restrict__ basis)
{
    constexpr int NM = NQ - 1, KS = (NQ + 3) / 4, MT = (NM + 7) / 8;
    for (int idx = threadIdx.x; idx < MT * KS * 32; idx += THREADS)
    {
        const int l = idx & 31, t = idx >> 5, ks = t % KS, mt = t / KS;
        const int q = 8 * mt + (l >> 2), j = 4 * ks + (l & 3);
        frag[idx]   = (q < NM && j < NQ) ? basis[q * NQ + j] : 0.0;
    }
}

template <int NQ, int G, int WARPS, bool WEIGHTED> struct HexIprodMma
{
    static constexpr int NM = NQ - 1, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ, NM2 = NM * NM, NM3 = NM2 * NM;
    static constexpr int PW  = ipm_pad8(NM);        // p (and q, r) padded to the tile width
    static constexpr int DS1 = ipm_stride(PW);      // s1[(e,kz,j)][p]
    static constexpr int DG2 = ipm_stride(PW * PW); // s2[(e,kz)][q][p], stride between kz
    static constexpr int KS = (NQ + 3) / 4, NT = PW / 8;
    static constexpr int SLOT = (G * NQ3 + 1 + 3 + 1) / 2 * 2;
    // first intermediate, later the staged result s3[e][r][q][p] (whichever is larger)
    static constexpr int S1   = G * NQ2 * DS1 > G * PW * PW * PW ? G * NQ2 * DS1 : G * PW * PW * PW;
    static constexpr int S2   = G * NQ * DG2;
    static constexpr int WARP_D = (WEIGHTED ? 2 : 1) * SLOT + S1 + S2;
    static constexpr int FRAG = KS * NT * 32;
    static constexpr int BAR_BYTES = (WARPS * 8 + 15) / 16 * 16;
    static constexpr size_t SMEM = BAR_BYTES + (size_t)(3 * FRAG + WARPS * WARP_D) * sizeof(double);
    static_assert(G * PW * PW * PW <= S1, "the staged result must fit the first intermediate");
};

template <int NQ, int G, int WARPS, int MB, int NB, bool WEIGHTED>
__global__ void __launch_bounds__(WARPS * 32)
    iproduct_hex_mma_kernel(const double *__restrict__ basis0, const double *__restrict__ basis1,
                            const double *__restrict__ basis2, const double *__restrict__ in,
                            const double *__restrict__ wgt, double *__restrict__ out, unsigned nelmt, unsigned ngroups)
{
    using C = HexIprodMma<NQ, G, WARPS, WEIGHTED>;
    constexpr int NM = C::NM, PW = C::PW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    double *frag0  = reinterpret_cast<double *>(smem_raw + C::BAR_BYTES);
    double *frag1  = frag0 + C::FRAG;
    double *frag2  = frag1 + C::FRAG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *slot  = frag2 + C::FRAG + warp * C::WARP_D;
    double *wslot = slot + C::SLOT; // only when WEIGHTED
    double *s1    = slot + (WEIGHTED ? 2 : 1) * C::SLOT;
    double *s2    = s1 + C::S1;
    uint64_t *bar = bars + warp;

    ipm_fill_fragB<NQ, WARPS * 32>(frag0, basis0);
    ipm_fill_fragA<NQ, WARPS * 32>(frag1, basis1);
    ipm_fill_fragA<NQ, WARPS * 32>(frag2, basis2);
    if (lane == 0)
    {
        mbar_init(bar, WEIGHTED ? 2 : 1);
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned nw = gridDim.x * WARPS;
    unsigned g        = blockIdx.x * WARPS + warp;
    unsigned parity   = 0;
    // both slabs share the element-major layout, so they take the same path (bulk copy / plain loads)
    auto fetch = [&](unsigned gn) {
        bool by_bar = mma_fetch_group<G, C::NQ3>(slot, bar, in, gn, nelmt, lane);
        if (WEIGHTED)
        {
            const bool wb = mma_fetch_group<G, C::NQ3>(wslot, bar, wgt, gn, nelmt, lane);
            if (wb != by_bar) // one of the two went through plain loads: balance the barrier's arrival count
            {
                if (lane == 0)
                    mbar_arrive_expect_tx(bar, 0);
                by_bar = true;
            }
        }
        return by_bar;
    };
    bool by_bar = false;
    if (g < ngroups)
        by_bar = fetch(g);

    for (; g < ngroups; g += nw)
    {
        const size_t e0 = (size_t)g * G;
        const int ne    = (nelmt - e0 < (size_t)G) ? (int)(nelmt - e0) : G;
        const double *s_in = slot + ((reinterpret_cast<uintptr_t>(in + e0 * C::NQ3) & 15u) >> 3);
        const double *s_w  = WEIGHTED ? wslot + ((reinterpret_cast<uintptr_t>(wgt + e0 * C::NQ3) & 15u) >> 3) : s_in;
        if (by_bar)
        {
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        // direction 0: s1[(e,kz,j)][p] = sum_i (w*in)[(e,kz,j)][i] B0[p][i]
        ipm_pass_data<NQ, NM, C::DS1, MB, G * C::NQ2, WEIGHTED>(s_in, s_w, frag0, s1, ne * C::NQ2, lane);
        __syncwarp();
        if (g + nw < ngroups)
            by_bar = fetch(g + nw);
        // direction 1: s2[(e,kz)][q][p] = sum_j B1[q][j] s1[(e,kz)][j][p]
        ipm_pass_basis<NM, NQ, PW, C::DS1, C::DG2, PW, NB, G * NQ * PW>(s1, frag1, s2, ne * NQ * PW, lane);
        __syncwarp();
        // direction 2: s3[e][r][(q,p)] = sum_kz B2[r][kz] s2[e][kz][(q,p)]   (s3 aliases s1)
        ipm_pass_basis<NM, NQ, PW * PW, C::DG2, PW * PW * PW, PW * PW, NB, G * PW * PW>(s2, frag2, s1, ne * PW * PW,
                                                                                       lane);
        __syncwarp();
        // staged result -> out[e][r][q][p], unpadded and contiguous
        double *gout = out + e0 * C::NM3;
        for (int idx = lane; idx < ne * C::NM3; idx += 32)
        {
            const int e = idx / C::NM3, rqp = idx - e * C::NM3, r = rqp / C::NM2, qp = rqp - r * C::NM2;
            const int q = qp / NM, p = qp - q * NM;
            st_stream(gout + idx, s1[((e * PW + r) * PW + q) * PW + p]);
        }
        __syncwarp(); // s1 / s2 are rewritten by the next group
    }
}

template <int NQ, int G, int WARPS, bool WEIGHTED> struct QuadIprodMma
{
    static constexpr int NM = NQ - 1, NQ2 = NQ * NQ, NM2 = NM * NM;
    static constexpr int PW  = ipm_pad8(NM);
    static constexpr int DS1 = ipm_stride(PW); // s1[(e,j)][p]
    static constexpr int KS = (NQ + 3) / 4, NT = PW / 8;
    static constexpr int SLOT = (G * NQ2 + 1 + 3 + 1) / 2 * 2;
    static constexpr int S1   = G * NQ * DS1;
    static constexpr int S2   = G * PW * PW;   // staged result s2[e][q][p]
    static constexpr int WARP_D = (WEIGHTED ? 2 : 1) * SLOT + S1 + S2;
    static constexpr int FRAG = KS * NT * 32;
    static constexpr int BAR_BYTES = (WARPS * 8 + 15) / 16 * 16;
    static constexpr size_t SMEM = BAR_BYTES + (size_t)(2 * FRAG + WARPS * WARP_D) * sizeof(double);
};

template <int NQ, int G, int WARPS, int MB, int NB, bool WEIGHTED>
__global__ void __launch_bounds__(WARPS * 32)
    iproduct_quad_mma_kernel(const double *__restrict__ basis0, const double *__restrict__ basis1,
                             const double *__restrict__ in, const double *__restrict__ wgt, double *__restrict__ out,
                             unsigned nelmt, unsigned ngroups)
{
    using C = QuadIprodMma<NQ, G, WARPS, WEIGHTED>;
    constexpr int NM = C::NM, PW = C::PW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    double *frag0  = reinterpret_cast<double *>(smem_raw + C::BAR_BYTES);
    double *frag1  = frag0 + C::FRAG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *slot  = frag1 + C::FRAG + warp * C::WARP_D;
    double *wslot = slot + C::SLOT;
    double *s1    = slot + (WEIGHTED ? 2 : 1) * C::SLOT;
    double *s2    = s1 + C::S1;
    uint64_t *bar = bars + warp;

    ipm_fill_fragB<NQ, WARPS * 32>(frag0, basis0);
    ipm_fill_fragA<NQ, WARPS * 32>(frag1, basis1);
    if (lane == 0)
    {
        mbar_init(bar, WEIGHTED ? 2 : 1);
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned nw = gridDim.x * WARPS;
    unsigned g        = blockIdx.x * WARPS + warp;
    unsigned parity   = 0;
    auto fetch = [&](unsigned gn) {
        bool by_bar = mma_fetch_group<G, C::NQ2>(slot, bar, in, gn, nelmt, lane);
        if (WEIGHTED)
        {
            const bool wb = mma_fetch_group<G, C::NQ2>(wslot, bar, wgt, gn, nelmt, lane);
            if (wb != by_bar)
            {
                if (lane == 0)
                    mbar_arrive_expect_tx(bar, 0);
                by_bar = true;
            }
        }
        return by_bar;
    };
    bool by_bar = false;
    if (g < ngroups)
        by_bar = fetch(g);

    for (; g < ngroups; g += nw)
    {
        const size_t e0 = (size_t)g * G;
        const int ne    = (nelmt - e0 < (size_t)G) ? (int)(nelmt - e0) : G;
        const double *s_in = slot + ((reinterpret_cast<uintptr_t>(in + e0 * C::NQ2) & 15u) >> 3);
        const double *s_w  = WEIGHTED ? wslot + ((reinterpret_cast<uintptr_t>(wgt + e0 * C::NQ2) & 15u) >> 3) : s_in;
        if (by_bar)
        {
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        // direction 0: s1[(e,j)][p] = sum_i (w*in)[(e,j)][i] B0[p][i]
        ipm_pass_data<NQ, NM, C::DS1, MB, G * NQ, WEIGHTED>(s_in, s_w, frag0, s1, ne * NQ, lane);
        __syncwarp();
        if (g + nw < ngroups)
            by_bar = fetch(g + nw);
        // direction 1: s2[e][q][p] = sum_j B1[q][j] s1[e][j][p]
        ipm_pass_basis<NM, NQ, PW, C::DS1, PW * PW, PW, NB, G * PW>(s1, frag1, s2, ne * PW, lane);
        __syncwarp();
        double *gout = out + e0 * C::NM2;
        for (int idx = lane; idx < ne * C::NM2; idx += 32)
        {
            const int e = idx / C::NM2, qp = idx - e * C::NM2, q = qp / NM, p = qp - q * NM;
            st_stream(gout + idx, s2[(e * PW + q) * PW + p]);
        }
        __syncwarp();
    }
}

} // namespace b200fe

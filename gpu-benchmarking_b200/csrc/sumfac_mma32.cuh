// sumfac_mma32.cuh -- "mma" back-end for FP32: sum-factorisation on the warp-level tensor-core
// path with the 3xTF32 split (mma.sync.m16n8k8.tf32, FP32 accumulate) for the large-nq quads.
//
// FP32 quads with nq >= 12 need 37-103 TFLOP/s at the HBM roofline; the FFMA pipe peaks at 72
// TFLOP/s on this B200 and reaches 28-40 inside the rows kernels (profiles/r01_ubench_fp_pipe.txt).
// HMMA.1688.F32.TF32 sustains 278 TFLOP/s (profiles/r01_ubench_tf32.txt).  A plain TF32 product
// keeps 11 mantissa bits (relative error 5e-4): not acceptable against the 1e-5 bar.  Splitting both
// operands x = x_hi + x_lo with x_hi = tf32(x) and accumulating
//        a_lo*b_hi + a_hi*b_lo + a_hi*b_hi          (a_lo*b_lo ~ 2^-22 |ab| is dropped)
// in FP32 restores ~2^-21 relative accuracy per product at three tensor-core instructions per
// tile, i.e. 93 TFLOP/s of FP32-equivalent peak with 1/16 of the issue slots FFMA would need.
// Results agree with the reference's FFMA chain to rounding (measured max error relative to the
// largest output ~1e-6, tests hold 1e-5), NOT bit for bit; the bit-exact FP32 path remains
// rows / pipe (b200fe_set_backend).
//
// Structure identical to the FP64 kernel in sumfac_mma.cuh: one warp owns a group of G elements,
// the group's slab arrives by one bulk (TMA) copy on the warp's mbarrier, direction 0 uses the
// data as the A operand (M = flattened rows), direction 1 uses the transposed basis as A and the
// intermediate as B (N = flattened (e, i)), every lane ends with out[e][j][i..i+1] (8-byte store).
#pragma once

#include "sumfac_mma.cuh"

namespace b200fe
{

__device__ __forceinline__ void hmma1688_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ unsigned tf32_hi(float x)
{
    unsigned h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    return h;
}
// x = hi + lo exactly, both parts rounded to tf32 (used once, for the basis fragments)
__device__ __forceinline__ void tf32_split(float x, unsigned &hi, unsigned &lo)
{
    hi = tf32_hi(x);
    lo = tf32_hi(x - __uint_as_float(hi));
}
// the per-value split of the data stream in 3 instructions: cvt.rna.tf32 expands to ~5 ALU instructions on
// sm_100 (profiles/r01_ncu_quad16_f32_mma.txt: the conversions were 3/4 of the instruction stream), so hi is
// rounded to nearest (ties away) by an integer add + mask on the sign-magnitude bit pattern and lo = x - hi
// (exact) is handed to the tensor core as is, which truncates it to tf32 (error <= 2^-21 |x|).  A NaN input
// still yields NaN through lo; values within half a tf32 ulp of FLT_MAX round to Inf like cvt.rna would.
__device__ __forceinline__ void tf32_split_fast(float x, unsigned &hi, unsigned &lo)
{
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

template <int NQ, int G, int WARPS, int MB0, int NB1> struct QuadMma32
{
    static constexpr int NM    = NQ - 1;
    static constexpr int NM2   = NM * NM;
    static constexpr int NQ2   = NQ * NQ;
    static constexpr int KS    = (NM + 7) / 8;   // k steps of 8
    static constexpr int NT0   = (NQ + 7) / 8;   // direction 0: n tiles over i
    static constexpr int MT1   = (NQ + 15) / 16; // direction 1: m tiles over j
    static constexpr int ROWS0 = G * NM;
    static constexpr int MT0   = (ROWS0 + 15) / 16;
    static constexpr int COLS1 = G * NQ;
    static constexpr int NT1   = (COLS1 + 7) / 8;
    // row stride of the intermediate (even); == 8 (mod 16) where that costs little, so that the rows a store /
    // a direction-1 B fragment touches fall into distinct banks
    static constexpr int S     = (NQ % 16 == 0) ? NQ + 8 : NQ;
    static constexpr int SLOT  = (G * NM2 + 3 + 7 + 3) / 4 * 4; // +3: window offset, +7: k over-read of the last row
    static constexpr int MID   = (G * NM * S + 3) / 4 * 4;
    static constexpr int WARP_F = SLOT + MID;    // floats per warp (multiple of 4)
    static constexpr int FRAG0 = KS * NT0 * 32 * 4;     // [ks][nt][lane][hi b0, hi b1, lo b0, lo b1]
    static constexpr int FRAG1 = MT1 * KS * 2 * 32 * 4; // [mt][ks][hi|lo][lane][a0..a3]
    static constexpr int BAR_BYTES = (WARPS * 8 + 15) / 16 * 16;
    static constexpr size_t SMEM = BAR_BYTES + (size_t)(FRAG0 + FRAG1 + WARPS * WARP_F) * sizeof(float);
    static_assert(NQ % 2 == 0, "the mma back-end pairs outputs along i");
    static_assert(COLS1 % 8 == 0, "G * nq must fill whole column tiles");
};

// group fetch, FP32 twin of mma_fetch_group: bulk copy of the enclosing 16-byte aligned window
template <int G, int NM2>
__device__ __forceinline__ bool mma32_fetch_group(float *slot, uint64_t *bar, const float *__restrict__ in, unsigned gn,
                                                  unsigned nelmt, int lane)
{
    const size_t e0      = (size_t)gn * G;
    const unsigned ne    = (nelmt - e0 < (size_t)G) ? (unsigned)(nelmt - e0) : (unsigned)G;
    const unsigned count = ne * (unsigned)NM2;
    const float *src     = in + e0 * NM2;
    const unsigned off   = (unsigned)((reinterpret_cast<uintptr_t>(src) & 15u) >> 2);
    const float *wsrc    = src - off;
    const unsigned bytes = ((off + count) * 4u + 15u) & ~15u;
    const bool fits      = (wsrc >= in) && (reinterpret_cast<const char *>(wsrc) + bytes <=
                                       reinterpret_cast<const char *>(in + (size_t)nelmt * NM2));
    if (fits)
    {
        if (lane == 0)
        {
            fence_proxy_async();
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

template <int NQ, int G, int WARPS, int MB0, int NB1, bool SUMSQ = false>
__global__ void __launch_bounds__(WARPS * 32)
    bwdtrans_quad_mma32_kernel(const float *__restrict__ basis0, const float *__restrict__ basis1,
                               const float *__restrict__ in, float *__restrict__ out, unsigned nelmt, unsigned ngroups,
                               int out_vec, double *__restrict__ partials)
{
    using C = QuadMma32<NQ, G, WARPS, MB0, NB1>;
    constexpr int NM = C::NM, KS = C::KS, NT0 = C::NT0, MT1 = C::MT1, MT0 = C::MT0, NT1 = C::NT1, S = C::S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bars  = reinterpret_cast<uint64_t *>(smem_raw);
    unsigned *frag0 = reinterpret_cast<unsigned *>(smem_raw + C::BAR_BYTES);
    unsigned *frag1 = frag0 + C::FRAG0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    float *slot   = reinterpret_cast<float *>(frag1 + C::FRAG1) + warp * C::WARP_F;
    float *mid    = slot + C::SLOT;
    uint64_t *bar = bars + warp;

    // basis matrices in fragment order, split into tf32 hi / lo parts, zero padded
    for (int idx = threadIdx.x; idx < C::FRAG0 / 2; idx += WARPS * 32)
    {
        const int l = idx & 31, which = (idx >> 5) & 1, tile = idx >> 6, nt = tile % NT0, ks = tile / NT0;
        const int p = 8 * ks + (l & 3) + 4 * which, i = 8 * nt + (l >> 2);
        const float v = (p < NM && i < NQ) ? basis0[p * NQ + i] : 0.0f;
        unsigned hi, lo;
        tf32_split(v, hi, lo);
        frag0[(tile * 32 + l) * 4 + which]     = hi; // one 16-byte load per lane and tile: {hi b0, hi b1, lo b0, lo b1}
        frag0[(tile * 32 + l) * 4 + 2 + which] = lo;
    }
    for (int idx = threadIdx.x; idx < C::FRAG1 / 2; idx += WARPS * 32)
    {
        const int l = idx & 31, which = (idx >> 5) & 3, tile = idx >> 7, ks = tile % KS, mt = tile / KS;
        const int j = 16 * mt + (l >> 2) + 8 * (which & 1), q = 8 * ks + (l & 3) + 4 * (which >> 1);
        const float v = (j < NQ && q < NM) ? basis1[q * NQ + j] : 0.0f;
        unsigned hi, lo;
        tf32_split(v, hi, lo);
        frag1[((tile * 2 + 0) * 32 + l) * 4 + which] = hi; // two 16-byte loads per lane and tile: hi a0..a3, lo a0..a3
        frag1[((tile * 2 + 1) * 32 + l) * 4 + which] = lo;
    }
    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned nw = gridDim.x * WARPS;
    unsigned grp      = blockIdx.x * WARPS + warp;
    unsigned parity   = 0;
    double ss         = 0.0; // fused checksum (SUMSQ), accumulated in double like b200fe_sumsq_f32
    bool by_bar       = false;
    if (grp < ngroups)
        by_bar = mma32_fetch_group<G, C::NM2>(slot, bar, in, grp, nelmt, lane);

    for (; grp < ngroups; grp += nw)
    {
        const size_t e0 = (size_t)grp * G;
        const int ne    = (nelmt - e0 < (size_t)G) ? (int)(nelmt - e0) : G;
        const float *s_in = slot + ((reinterpret_cast<uintptr_t>(in + e0 * C::NM2) & 15u) >> 2);
        if (by_bar)
        {
            mbar_wait(bar, parity);
            parity ^= 1u;
        }

        // ---- direction 0: mid[(e,q)][i] = sum_p in[(e,q)][p] B0[p][i];  A = data ---------------------
        // Rows past the group's last one (tile padding, ragged last group) read whatever lies there inside
        // this warp's region: rows are independent and those outputs are not stored.
        {
            const int nrows    = ne * NM;
            const float *abase = s_in + g * NM + t;
#pragma unroll
            for (int mt = 0; mt < MT0; mt += MB0)
            {
                float acc[MB0][NT0][4];
#pragma unroll
                for (int m = 0; m < MB0; ++m)
#pragma unroll
                    for (int n = 0; n < NT0; ++n)
                        acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.0f;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                {
                    unsigned ahi[MB0][4], alo[MB0][4];
#pragma unroll
                    for (int m = 0; m < MB0; ++m)
                        if (mt + m < MT0)
                        {
                            const float *ap = abase + (mt + m) * 16 * NM + 8 * ks;
                            float v[4]      = {ap[0], ap[8 * NM], ap[4], ap[8 * NM + 4]};
                            if (8 * ks + 7 >= NM) // k padding must not contribute whatever lies there
                            {
                                if (8 * ks + t >= NM)
                                    v[0] = v[1] = 0.0f;
                                if (8 * ks + t + 4 >= NM)
                                    v[2] = v[3] = 0.0f;
                            }
#pragma unroll
                            for (int x = 0; x < 4; ++x)
                                tf32_split_fast(v[x], ahi[m][x], alo[m][x]);
                        }
#pragma unroll
                    for (int n = 0; n < NT0; ++n)
                    {
                        const uint4 fb = *reinterpret_cast<const uint4 *>(frag0 + ((ks * NT0 + n) * 32 + lane) * 4);
                        const unsigned bhi[2] = {fb.x, fb.y}, blo[2] = {fb.z, fb.w};
#pragma unroll
                        for (int m = 0; m < MB0; ++m)
                            if (mt + m < MT0)
                            {
                                hmma1688_tf32(acc[m][n], alo[m], bhi);
                                hmma1688_tf32(acc[m][n], ahi[m], blo);
                                hmma1688_tf32(acc[m][n], ahi[m], bhi);
                            }
                    }
                }
#pragma unroll
                for (int m = 0; m < MB0; ++m)
                    if (mt + m < MT0)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                        {
                            const int row = (mt + m) * 16 + g + 8 * h;
                            if (row < nrows)
                            {
#pragma unroll
                                for (int n = 0; n < NT0; ++n)
                                    if (8 * n + 8 <= NQ || 8 * n + 2 * t < NQ)
                                        *reinterpret_cast<float2 *>(mid + row * S + 8 * n + 2 * t) =
                                            make_float2(acc[m][n][2 * h], acc[m][n][2 * h + 1]);
                            }
                        }
            }
        }
        __syncwarp();
        if (grp + nw < ngroups)
            by_bar = mma32_fetch_group<G, C::NM2>(slot, bar, in, grp + nw, nelmt, lane);

        // ---- direction 1: out[e][j][i] = sum_q B1[q][j] mid[e][q][i];  A = transposed basis ----------
        {
            const int ncols    = ne * NQ;
            float *gout        = out + e0 * C::NQ2;
            const float *bbase = mid + t * S + g;
#pragma unroll
            for (int nt = 0; nt < NT1; nt += NB1)
            {
                float acc[MT1][NB1][4];
#pragma unroll
                for (int m = 0; m < MT1; ++m)
#pragma unroll
                    for (int n = 0; n < NB1; ++n)
                        acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.0f;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                {
                    unsigned bhi[NB1][2], blo[NB1][2];
#pragma unroll
                    for (int n = 0; n < NB1; ++n)
                        if (nt + n < NT1)
                        {
                            const int e = ((nt + n) * 8) / NQ, w0 = ((nt + n) * 8) % NQ;
                            // columns w0 + g >= nq belong to the next element: its rows start nm*S later
                            const float *bp = ((w0 + 8 > NQ && g >= NQ - w0) ? bbase + (NM * S - NQ) : bbase) +
                                              (e * NM + 8 * ks) * S + w0;
                            float v[2];
                            v[0] = (8 * ks + 7 < NM || 8 * ks + t < NM) ? bp[0] : 0.0f;         // rows >= nm do not exist
                            v[1] = (8 * ks + 7 < NM || 8 * ks + t + 4 < NM) ? bp[4 * S] : 0.0f;
                            tf32_split_fast(v[0], bhi[n][0], blo[n][0]);
                            tf32_split_fast(v[1], bhi[n][1], blo[n][1]);
                        }
#pragma unroll
                    for (int m = 0; m < MT1; ++m)
                    {
                        const uint4 *fa = reinterpret_cast<const uint4 *>(frag1 + ((m * KS + ks) * 2 * 32 + lane) * 4);
                        const uint4 fh = fa[0], fl = fa[32];
                        const unsigned ahi[4] = {fh.x, fh.y, fh.z, fh.w};
                        const unsigned alo[4] = {fl.x, fl.y, fl.z, fl.w};
#pragma unroll
                        for (int n = 0; n < NB1; ++n)
                            if (nt + n < NT1)
                            {
                                hmma1688_tf32(acc[m][n], alo, bhi[n]);
                                hmma1688_tf32(acc[m][n], ahi, blo[n]);
                                hmma1688_tf32(acc[m][n], ahi, bhi[n]);
                            }
                    }
                }
#pragma unroll
                for (int n = 0; n < NB1; ++n)
                    if (nt + n < NT1)
                    {
                        const int col = (nt + n) * 8 + 2 * t;
                        if (col < ncols)
                        {
                            const int e = col / NQ, w = col - e * NQ;
                            float *op = gout + (size_t)e * C::NQ2 + w;
#pragma unroll
                            for (int m = 0; m < MT1; ++m)
#pragma unroll
                                for (int h = 0; h < 2; ++h)
                                {
                                    const int j = 16 * m + g + 8 * h;
                                    if (j < NQ)
                                    {
                                        if (SUMSQ)
                                        {
                                            const double a0 = acc[m][n][2 * h], a1 = acc[m][n][2 * h + 1];
                                            ss = fmadd(a1, a1, fmadd(a0, a0, ss));
                                        }
                                        if (out_vec)
                                            st_stream(reinterpret_cast<float2 *>(op + j * NQ),
                                                      make_float2(acc[m][n][2 * h], acc[m][n][2 * h + 1]));
                                        else
                                        {
                                            st_stream(op + j * NQ, acc[m][n][2 * h]);
                                            st_stream(op + j * NQ + 1, acc[m][n][2 * h + 1]);
                                        }
                                    }
                                }
                        }
                    }
            }
        }
        __syncwarp(); // mid is rewritten by the next group's direction 0
    }
    if (SUMSQ)
        mma_store_partial(ss, partials, blockIdx.x * WARPS + warp, lane);
}

} // namespace b200fe

// sumfac_coamma.cuh -- "coa-mma" back-end: FP64 quads in the warp-interleaved layout of the reference's "Coales"
// kernel (benchmark04.cc:78-147)
//     x[(e/32)*32*len + 32*idx + e%32]
// on the FP64 tensor cores (mma.sync.m8n8k4.f64 = SASS DMMA.8x8x4) with M = ELEMENTS.
//
// At nq = 32 the operator is compute-bound (7.9 flop/B: 51 TFLOP/s at the HBM roofline against 37 TFLOP/s of FP64
// peak) and the lanes kernel that served this layout -- one row of 31 values per thread, DFMA with a constant-bank
// operand in a rolled loop -- reaches 0.29 of the roofline (profiles/r02_ncu_quad32_f64_coa.txt: FP64 pipe 42 %,
// mio_throttle).  The interleaved layout is itself a GEMM operand: for a fixed row q
//     t1[q][i][e] = sum_p in[q][p][e] * B0[p][i]          A[m = e][k = p]  (m contiguous),  B = B0,  N = i
//     out[j][i][e] = sum_q t1[q][i][e] * B1[q][j]         A[m = e][k = q],                  B = B1,  N = j
// so a tile of 8 consecutive elements of an interleave group is an m8 tile as it lies in memory, K = nm padded to
// nq with zeros in the basis fragments (3 % waste at nq = 32, none in N), and nothing is ever transposed.
//
//   gather  the tile [idx][8 e] (64-byte runs, 256 bytes apart) by 16-byte cp.async copies into ONE region of
//           nm rows x nq units (a unit = one index of the 8 elements = 64 bytes): row q of the input (nm units)
//           at units [nq*q, nq*q + nm)
//   dir 0   a warp takes rows q, q + WARPS, ...: KS A fragments from its row (lanes (e, p)), KS x NT DMMAs against the
//           B0 fragments held in registers, and t1[q][.][e] (nq units) goes back IN PLACE over the row it came from --
//           the row belongs to that warp alone, so direction 0 needs no barrier at all
//   dir 1   a warp takes columns i: A fragments from t1 (lanes (e, q)), B1 fragments in registers, results straight
//           to global memory: each store instruction writes four whole 64-byte runs
// 63.5 KB per tile + 16 KB of basis fragments: two CTAs per SM whose gather / contract phases overlap (a first version
// with a two-slot ring and one CTA per SM left the tensor pipe idle through every gather issue, fragment reload and
// barrier: 65 % DMMA utilisation, 0.45 of the roofline).
// Bank conflicts.  An 8-byte access is served a half warp at a time: fragment rows 0-3 (4 elements = 32 bytes) x the 4
// fragment columns, which therefore have to fall into the four 32-byte quarters of a 128-byte bank window.  The tile's
// unit is 64 bytes (index u, 8 elements), so consecutive u alternate between the two halves and a two-bit swizzle does
// the rest (first version without it: 43 % of the shared-memory wavefronts were conflict replays):
//   input  u = 32 q + p: element e at e ^ 4*(u>>1 & 1)                 (A loads of direction 0: 4 consecutive u)
//   t1     u = 32 q + i stored at unit u ^ ((u>>1 ^ u>>5) & 1), element e ^ 4*((u>>2 ^ u>>6) & 1)
//          (C stores of direction 0: 4 units two apart; A loads of direction 1: 4 units 32 apart)
// The gather writes the input swizzled (16-byte chunk `part` of an index goes to part ^ 2*(u>>1 & 1)).
// DMMA.8x8x4 on sm_100 accumulates its four products in k order with fused multiply-adds (sumfac_mma.cuh), and the
// k steps are issued in ascending order onto one accumulator: the reference's summation order, bit-identical.
#pragma once

#include "sumfac_coapipe.cuh"
#include "sumfac_mma.cuh"

namespace b200fe
{

template <int NQ, int WARPS> struct QuadCoaMma
{
    static_assert(NQ % 8 == 0, "whole n tiles");
    static constexpr int EL = 8, NM = NQ - 1, NM2 = NM * NM, NQ2 = NQ * NQ;
    static constexpr int KS = (NM + 3) / 4, NT = NQ / 8;
    static constexpr int S1   = NM * NQ * EL;   // the region: input rows pitched to nq units, then t1 in place (doubles)
    static constexpr int FRAG = KS * NT * 32;   // one basis matrix in fragment order
    static constexpr size_t SMEM = (size_t)(S1 + 2 * FRAG) * sizeof(double);
    static constexpr int THREADS = WARPS * 32;
    static constexpr int PER = 32 / EL;
};

// RB rows (columns) per warp at once: RB * NT independent accumulator chains in flight
template <int NQ, int WARPS, int RB = 1>
__global__ void __launch_bounds__(WARPS * 32, 2)
    bwdtrans_quad_coamma_kernel(const double *__restrict__ b0, const double *__restrict__ b1,
                                const double *__restrict__ in, double *__restrict__ out, unsigned ntiles)
{
    using C = QuadCoaMma<NQ, WARPS>;
    constexpr int NM = C::NM, NM2 = C::NM2, KS = C::KS, NT = C::NT, EL = C::EL, PER = C::PER;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s1    = reinterpret_cast<double *>(smem_raw);
    double *fb0   = s1 + C::S1; // (also what the zero-selected reads past the end of t1 land in)
    double *fb1   = fb0 + C::FRAG;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, c = lane & 3; // fragment row (element) / column

    // basis fragments: B[k = 4 ks + c][n = 8 nt + r], zero where k is padding
    for (int f = tid; f < C::FRAG; f += C::THREADS)
    {
        const int l = f & 31, t = f >> 5, ks = t / NT, nt = t - ks * NT;
        const int k = 4 * ks + (l & 3), n = 8 * nt + (l >> 2);
        fb0[f] = k < NM ? b0[k * NQ + n] : 0.0;
        fb1[f] = k < NM ? b1[k * NQ + n] : 0.0;
    }

    auto issue = [&](unsigned tile) {
        const unsigned group = tile / PER, l0 = (tile % PER) * EL;
        const double *g      = in + (size_t)group * 32 * NM2 + l0;
#pragma unroll 4
        for (int ch = tid; ch < NM2 * 4; ch += C::THREADS)
        {
            const int idx = ch >> 2, part = ch & 3, q = idx / NM, u = idx + q; // u = nq*q + p
            cp_async16(s1 + u * EL + 2 * (part ^ (u & 2)), g + (size_t)idx * 32 + part * 2);
        }
        cp_async_commit();
    };

    const bool kpad = 4 * (KS - 1) + c >= NM; // this lane's column of the last k step is padding

#pragma unroll 1
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
        __syncthreads(); // every warp has left direction 1 of the previous tile (and the fragments are built)
        issue(tile);
        cp_async_wait<0>();
        __syncthreads();

        double b[KS][NT];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int n = 0; n < NT; ++n)
                b[ks][n] = fb0[(ks * NT + n) * 32 + lane];

        // direction 0: rows q of the 8 elements
#pragma unroll 1
        for (int q0 = warp; q0 < NM; q0 += RB * WARPS)
        {
            double a[RB][KS], acc[RB][NT][2];
#pragma unroll
            for (int k = 0; k < RB; ++k)
            {
                const int q = q0 + k * WARPS < NM ? q0 + k * WARPS : q0; // (clamped: recomputes row q0, stores nothing)
                // u = nq*q + 4 ks + c: bit 1 of u is that of c
                const double *ap = s1 + (q * NQ + c) * EL + (r ^ (2 * (c & 2)));
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                    a[k][ks] = ap[4 * ks * EL]; // the padded column reads the row's spare unit (never written by the gather)
                if (kpad)
                    a[k][KS - 1] = 0.0; // ... which must not contribute, whatever it is (Inf, NaN)
#pragma unroll
                for (int n = 0; n < NT; ++n)
                    acc[k][n][0] = acc[k][n][1] = 0.0;
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int k = 0; k < RB; ++k)
#pragma unroll
                    for (int n = 0; n < NT; ++n)
                        dmma884_ordered(acc[k][n], a[k][ks], b[ks][n]);
#pragma unroll
            for (int k = 0; k < RB; ++k)
            {
                const int q = q0 + k * WARPS;
                if (q < NM)
                {
                    // u = 32 q + 8 n + 2 c + h: bit 1 = c & 1, bit 2 = c >> 1, bit 5 = q & 1, bit 6 = q >> 1 & 1
                    const int sw = (c ^ q) & 1, er = r ^ (4 * (((c >> 1) ^ (q >> 1)) & 1));
                    double *d0   = s1 + (q * NQ + 2 * c + sw) * EL + er;       // h = 0
                    double *d1   = s1 + (q * NQ + 2 * c + (sw ^ 1)) * EL + er; // h = 1
#pragma unroll
                    for (int n = 0; n < NT; ++n)
                    {
                        d0[8 * n * EL] = acc[k][n][0];
                        d1[8 * n * EL] = acc[k][n][1];
                    }
                }
            }
        }
        __syncthreads();

#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int n = 0; n < NT; ++n)
                b[ks][n] = fb1[(ks * NT + n) * 32 + lane];

        // direction 1: column i of the 8 elements, straight to global
        const unsigned group = tile / PER, l0 = (tile % PER) * EL;
        double *gout         = out + (size_t)group * 32 * C::NQ2 + l0 + r;
#pragma unroll 1
        for (int i0 = warp; i0 < NQ; i0 += RB * WARPS)
        {
            double a[RB][KS], acc[RB][NT][2];
#pragma unroll
            for (int k = 0; k < RB; ++k)
            {
                const int i = i0 + k * WARPS < NQ ? i0 + k * WARPS : i0;
                // u = 32 (4 ks + c) + i: bit 1 = i >> 1, bit 2 = i >> 2, bit 5 = c & 1, bit 6 = c >> 1 -- none depends on ks
                const double *ap = s1 + (c * NQ + (i ^ (((i >> 1) ^ c) & 1))) * EL + (r ^ (4 * (((i >> 2) ^ (c >> 1)) & 1)));
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                    a[k][ks] = ap[4 * ks * NQ * EL]; // q = nm (padding) reads the fragment area behind t1
                if (kpad)
                    a[k][KS - 1] = 0.0;
#pragma unroll
                for (int n = 0; n < NT; ++n)
                    acc[k][n][0] = acc[k][n][1] = 0.0;
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int k = 0; k < RB; ++k)
#pragma unroll
                    for (int n = 0; n < NT; ++n)
                        dmma884_ordered(acc[k][n], a[k][ks], b[ks][n]);
#pragma unroll
            for (int k = 0; k < RB; ++k)
            {
                const int i = i0 + k * WARPS;
                if (i < NQ)
                {
                    double *dst = gout + (size_t)32 * ((2 * c) * NQ + i); // j = 8 n + 2 c + h
#pragma unroll
                    for (int n = 0; n < NT; ++n)
                    {
                        st_stream(dst + (size_t)32 * NQ * (8 * n), acc[k][n][0]);
                        st_stream(dst + (size_t)32 * NQ * (8 * n + 1), acc[k][n][1]);
                    }
                }
            }
        }
    }
}

} // namespace b200fe

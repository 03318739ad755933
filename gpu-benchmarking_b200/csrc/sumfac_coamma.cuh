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
//           nq rows x nq units (a unit = one index of the 8 elements = 64 bytes): row q of the input (nm units)
//           at units [nq*q, nq*q + nm)
//   dir 0   a warp takes rows q, q + WARPS, ...: KS A fragments from its row (lanes (e, p)), KS x NT DMMAs against the
//           B0 fragments, and t1[q][.][e] (nq units) goes back IN PLACE over the row it came from -- the row belongs
//           to that warp alone, so direction 0 needs no barrier at all
//   dir 1   a warp takes columns i: A fragments from t1 (lanes (e, q)), B1 fragments, results straight to global
//           memory: each store instruction writes four whole 64-byte runs
// Both basis matrices live in registers, as fragments, for the whole kernel, and every warp runs its rows (columns) as
// a three-stage software pipeline -- A fragments of row k+1 loaded, DMMAs of row k issued, results of row k-1 stored --
// on two alternating fragment / accumulator sets, so the tensor pipe sees one uninterrupted DMMA stream per warp.
// 255 registers: 4 warps per CTA, two CTAs per SM (64 KB each) whose gather / contract phases overlap.
// Measured at 65 536 elements (tools/tune/lanes_probe.cu, profiles/r02_coa_probe.csv): 0.55 of the HBM roofline =
// 29 TFLOP/s = 79 % of the FP64 tensor peak (the ceiling at that peak is 0.70; the element-major DMMA kernel holds
// 0.60; the lanes kernel this replaces 0.29).  Steps on the way: two-slot ring + fragments reloaded from shared
// memory at every phase change, one CTA per SM: 0.45 (tensor pipe idle through every gather issue, reload and barrier);
// the swizzles below: 0.46; in-place rows, two CTAs per SM: 0.50; registers + pipeline: 0.55.  Tried and dropped: the
// gather by tiled TMA (one box of {8, nm, 1} per row q; the copy engine cannot apply the input swizzle, so the direction-0
// loads replay 2-way): 0.42 -- and at 255 registers the kernel is sensitive to anything that costs one more (the same
// source with an unused mbarrier pointer spilled 16 bytes and fell to 0.46).
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
#include "sumfac_mma32.cuh"

namespace b200fe
{

template <int NQ, int WARPS> struct QuadCoaMma
{
    static_assert(NQ % 8 == 0, "whole n tiles");
    static constexpr int EL = 8, NM = NQ - 1, NM2 = NM * NM, NQ2 = NQ * NQ;
    static constexpr int KS = (NM + 3) / 4, NT = NQ / 8;
    static constexpr int S1 = NQ * NQ * EL; // the region: nm rows of nq units + one spare row (the zero-selected k padding of direction 1 reads it)
    static constexpr size_t SMEM = (size_t)S1 * sizeof(double);
    static constexpr int THREADS = WARPS * 32;
    static constexpr int PER = 32 / EL;
    static constexpr int ROWS = (NM + WARPS - 1) / WARPS, COLS = (NQ + WARPS - 1) / WARPS; // per warp
};

template <int NQ, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)
    bwdtrans_quad_coamma_kernel(const double *__restrict__ b0, const double *__restrict__ b1,
                                const double *__restrict__ in, double *__restrict__ out, unsigned ntiles)
{
    using C = QuadCoaMma<NQ, WARPS>;
    constexpr int NM = C::NM, NM2 = C::NM2, KS = C::KS, NT = C::NT, EL = C::EL, PER = C::PER;
    extern __shared__ __align__(128) unsigned char smem_raw128[]; // (own name: the TU also declares smem_raw with 16-byte alignment)
    double *s1 = reinterpret_cast<double *>(smem_raw128);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, c = lane & 3; // fragment row (element) / column

    // basis fragments B[k = 4 ks + c][n = 8 nt + r], zero where k is padding: in registers from here on
    double f0[KS][NT], f1[KS][NT];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int n = 0; n < NT; ++n)
        {
            const int k = 4 * ks + c;
            f0[ks][n]   = k < NM ? b0[k * NQ + 8 * n + r] : 0.0;
            f1[ks][n]   = k < NM ? b1[k * NQ + 8 * n + r] : 0.0;
        }
    const bool kpad = 4 * (KS - 1) + c >= NM; // this lane's column of the last k step is padding

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
    auto load_a = [&](double (&a)[KS], const double *ap, int stride) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
            a[ks] = ap[ks * stride];
        if (kpad)
            a[KS - 1] = 0.0; // the padded column must not contribute, whatever lies there (Inf, NaN)
    };
    auto mma_row = [&](double (&acc)[NT][2], const double (&a)[KS], const double (&f)[KS][NT]) {
#pragma unroll
        for (int n = 0; n < NT; ++n)
            acc[n][0] = acc[n][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int n = 0; n < NT; ++n)
                dmma884_ordered(acc[n], a[ks], f[ks][n]);
    };
    // A fragments of row q (direction 0): u = nq*q + 4 ks + c, bit 1 of u is that of c
    auto row_ptr = [&](int q) { return s1 + (q * NQ + c) * EL + (r ^ (2 * (c & 2))); };
    // t1[q][i = 8 n + 2 c + h]: u = 32 q + i -- bit 1 = c & 1, bit 2 = c >> 1, bit 5 = q & 1, bit 6 = q >> 1 & 1
    auto store_row = [&](int q, const double (&acc)[NT][2]) {
        const int sw = (c ^ q) & 1, er = r ^ (4 * (((c >> 1) ^ (q >> 1)) & 1));
        double *d0   = s1 + (q * NQ + 2 * c + sw) * EL + er;
        double *d1   = s1 + (q * NQ + 2 * c + (sw ^ 1)) * EL + er;
#pragma unroll
        for (int n = 0; n < NT; ++n)
        {
            d0[8 * n * EL] = acc[n][0];
            d1[8 * n * EL] = acc[n][1];
        }
    };
    // A fragments of column i (direction 1): u = 32 (4 ks + c) + i -- bit 1 = i >> 1, bit 2 = i >> 2, bit 5 = c & 1, bit 6 = c >> 1
    auto col_ptr = [&](int i) {
        return s1 + (c * NQ + (i ^ (((i >> 1) ^ c) & 1))) * EL + (r ^ (4 * (((i >> 2) ^ (c >> 1)) & 1)));
    };

#pragma unroll 1
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
        __syncthreads(); // every warp has left direction 1 of the previous tile
        issue(tile);
        cp_async_wait<0>();
        __syncthreads();

        // direction 0: rows q = warp + k*WARPS, in place (a row belongs to one warp)
        {
            double a[2][KS], acc[2][NT][2];
            load_a(a[0], row_ptr(warp), 4 * EL);
#pragma unroll
            for (int k = 0; k < C::ROWS; ++k)
            {
                const int q = warp + k * WARPS;
                if (k + 1 < C::ROWS && q + WARPS < NM)
                    load_a(a[(k + 1) & 1], row_ptr(q + WARPS), 4 * EL);
                if (q < NM)
                    mma_row(acc[k & 1], a[k & 1], f0);
                if (k > 0)
                    store_row(q - WARPS, acc[(k - 1) & 1]);
            }
            if (warp + (C::ROWS - 1) * WARPS < NM)
                store_row(warp + (C::ROWS - 1) * WARPS, acc[(C::ROWS - 1) & 1]);
        }
        __syncthreads();

        // direction 1: columns i = warp + k*WARPS, straight to global
        {
            const unsigned group = tile / PER, l0 = (tile % PER) * EL;
            double *gout         = out + (size_t)group * 32 * C::NQ2 + l0 + r;
            auto store_col = [&](int i, const double (&acc)[NT][2]) {
                double *dst = gout + (size_t)32 * ((2 * c) * NQ + i); // j = 8 n + 2 c + h
#pragma unroll
                for (int n = 0; n < NT; ++n)
                {
                    st_stream(dst + (size_t)32 * NQ * (8 * n), acc[n][0]);
                    st_stream(dst + (size_t)32 * NQ * (8 * n + 1), acc[n][1]);
                }
            };
            double a[2][KS], acc[2][NT][2];
            load_a(a[0], col_ptr(warp), 4 * NQ * EL);
#pragma unroll
            for (int k = 0; k < C::COLS; ++k)
            {
                const int i = warp + k * WARPS;
                if (k + 1 < C::COLS && i + WARPS < NQ)
                    load_a(a[(k + 1) & 1], col_ptr(i + WARPS), 4 * NQ * EL);
                if (i < NQ)
                    mma_row(acc[k & 1], a[k & 1], f1);
                if (k > 0)
                    store_col(i - WARPS, acc[(k - 1) & 1]);
            }
            if (warp + (C::COLS - 1) * WARPS < NQ)
                store_col(warp + (C::COLS - 1) * WARPS, acc[(C::COLS - 1) & 1]);
        }
    }
}

// ---- FP32 twin: the same formulation on the warp-level TF32 tensor-core path with the 3xTF32 split ---------------
// (mma.sync.m16n8k8.tf32 = SASS HMMA.1688.F32.TF32, FP32 accumulate; split and error analysis: sumfac_mma32.cuh).
// FP32 nq = 32 needs 103 TFLOP/s at the HBM roofline against 72 TFLOP/s of FFMA peak: the lanes kernel (rolled FFMA2
// loop) holds 0.33 of the roofline; three HMMAs per product give 93 TFLOP/s of FP32-equivalent peak.
// M = 16 elements: a tile is 16 consecutive elements of an interleave group, the unit (one index of the tile) is again
// 64 bytes, so region, gather and the two-bit swizzles are those of the FP64 kernel with 8-float quarters.  A lane's A
// fragment is (e = g, g + 8) x (k = t, t + 4): four 4-byte loads per k step, each conflict-free across the warp.  The
// data is split on the fly (3 ALU instructions per value); the basis sits in a fragment-ordered area in shared memory
// ([ks][n][lane] x {b0, b1}, raw FP32: 8 KB for both matrices) and is loaded + split once per direction and tile.
// Measured (tools/tune/lanes_probe.cu): 0.49 of the roofline with 4 warps x 3 CTAs per SM (168 registers, 72 KB; 0.48 with
// two CTAs at 254 registers), HMMA pipe 59 % busy (profiles/r02_ncu_quad32_f32_coamma.txt); issuing two rows' chains
// interleaved instead of prefetching the next row was slower (0.45).
// Not bit-identical to the reference's FFMA chain: held to the component-wise 1e-5 bound of include/b200fe.h like the
// element-major tensor-core kernels (measured 7e-7 of the largest output); b200fe_set_backend("lanes") keeps the
// bit-exact kernel.
template <int NQ, int WARPS> struct QuadCoaMma32
{
    static_assert(NQ % 8 == 0, "whole n tiles and k steps");
    static constexpr int EL = 16, NM = NQ - 1, NM2 = NM * NM, NQ2 = NQ * NQ;
    static constexpr int KS = NQ / 8, NT = NQ / 8;
    static constexpr int S1   = NQ * NQ * EL;       // floats: nm rows of nq units + one spare row
    static constexpr int FRAG = KS * NT * 32 * 2;   // words per basis matrix: [ks][n][lane] x {b0, b1}, raw FP32
    static constexpr size_t SMEM = (size_t)(S1 + 2 * FRAG) * sizeof(float);
    static constexpr int THREADS = WARPS * 32;
    static constexpr int PER = 32 / EL;
    static constexpr int ROWS = (NM + WARPS - 1) / WARPS, COLS = (NQ + WARPS - 1) / WARPS; // per warp
};

template <int NQ, int WARPS, int MINB = 2>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    bwdtrans_quad_coamma32_kernel(const float *__restrict__ b0, const float *__restrict__ b1,
                                  const float *__restrict__ in, float *__restrict__ out, unsigned ntiles)
{
    using C = QuadCoaMma32<NQ, WARPS>;
    constexpr int NM = C::NM, NM2 = C::NM2, KS = C::KS, NT = C::NT, EL = C::EL, PER = C::PER;
    extern __shared__ __align__(128) unsigned char smem_raw128[]; // (own name: the TU also declares smem_raw with 16-byte alignment)
    float *s1     = reinterpret_cast<float *>(smem_raw128);
    float *fr0 = s1 + C::S1;
    float *fr1 = fr0 + C::FRAG;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    // basis fragments: b0 = B[k = 8 ks + t][n = 8 nt + g], b1 = B[k + 4][n]; zero where k is padding
    for (int f = tid; f < KS * NT * 32; f += C::THREADS)
    {
        const int l = f & 31, tt = f >> 5, ks = tt / NT, nt = tt - ks * NT;
        const int k = 8 * ks + (l & 3), n = 8 * nt + (l >> 2);
        *reinterpret_cast<float2 *>(fr0 + 2 * f) = make_float2(k < NM ? b0[k * NQ + n] : 0.f, k + 4 < NM ? b0[(k + 4) * NQ + n] : 0.f);
        *reinterpret_cast<float2 *>(fr1 + 2 * f) = make_float2(k < NM ? b1[k * NQ + n] : 0.f, k + 4 < NM ? b1[(k + 4) * NQ + n] : 0.f);
    }

    auto issue = [&](unsigned tile) {
        const unsigned group = tile / PER, l0 = (tile % PER) * EL;
        const float *gsrc    = in + (size_t)group * 32 * NM2 + l0;
#pragma unroll 4
        for (int ch = tid; ch < NM2 * 4; ch += C::THREADS)
        {
            const int idx = ch >> 2, part = ch & 3, q = idx / NM, u = idx + q; // u = nq*q + p
            cp_async16(s1 + u * EL + 4 * (part ^ (u & 2)), gsrc + (size_t)idx * 32 + part * 4);
        }
        cp_async_commit();
    };
    const bool kpad = t == 3; // this lane's second column (k = t + 4) of the last k step is padding (k = nm)
    // raw A fragments of one row / column: element offsets e0 (rows g) and e0 ^ 8 (rows g + 8), k = t and t + 4
    auto load_a = [&](float (&a)[KS][4], const float *ap, int e0, int stride) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
        {
            a[ks][0] = ap[ks * 8 * stride + e0];
            a[ks][1] = ap[ks * 8 * stride + (e0 ^ 8)];
            a[ks][2] = ap[(ks * 8 + 4) * stride + e0];
            a[ks][3] = ap[(ks * 8 + 4) * stride + (e0 ^ 8)];
        }
        if (kpad)
            a[KS - 1][2] = a[KS - 1][3] = 0.f; // must not contribute, whatever lies there (Inf, NaN)
    };
    auto mma_row = [&](float (&acc)[NT][4], const float (&a)[KS][4], const uint4 (&f)[KS][NT]) {
#pragma unroll
        for (int n = 0; n < NT; ++n)
            acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
        {
            unsigned hi[4], lo[4];
#pragma unroll
            for (int m = 0; m < 4; ++m)
                tf32_split_fast(a[ks][m], hi[m], lo[m]);
            // term-major: the NT accumulators are independent, the three terms of one accumulator are a chain
#pragma unroll
            for (int n = 0; n < NT; ++n)
            {
                const unsigned bh[2] = {f[ks][n].x, f[ks][n].y};
                hmma1688_tf32(acc[n], lo, bh);
            }
#pragma unroll
            for (int n = 0; n < NT; ++n)
            {
                const unsigned bl[2] = {f[ks][n].z, f[ks][n].w};
                hmma1688_tf32(acc[n], hi, bl);
            }
#pragma unroll
            for (int n = 0; n < NT; ++n)
            {
                const unsigned bh[2] = {f[ks][n].x, f[ks][n].y};
                hmma1688_tf32(acc[n], hi, bh);
            }
        }
    };
    // one 8-byte load per fragment, split here (exactly: both halves rounded to tf32) once per direction and tile
    auto load_frags = [&](uint4 (&f)[KS][NT], const float *fr) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int n = 0; n < NT; ++n)
            {
                const float2 v = *reinterpret_cast<const float2 *>(fr + 2 * ((ks * NT + n) * 32 + lane));
                tf32_split(v.x, f[ks][n].x, f[ks][n].z);
                tf32_split(v.y, f[ks][n].y, f[ks][n].w);
            }
    };
    // direction 0, row q: u = nq*q + 8 ks + t (+ 4): bit 1 of u is that of t
    const int e0row = g ^ (4 * (t & 2));
    // t1[q][i = 8 n + 2 t + h][e]: u = 32 q + i -- bit 1 = t & 1, bit 2 = t >> 1, bit 5 = q & 1, bit 6 = q >> 1 & 1
    auto store_row = [&](int q, const float (&acc)[NT][4]) {
        const int sw = (t ^ q) & 1, er = g ^ (8 * (((t >> 1) ^ (q >> 1)) & 1));
        float *d0    = s1 + (q * NQ + 2 * t + sw) * EL;       // h = 0
        float *d1    = s1 + (q * NQ + 2 * t + (sw ^ 1)) * EL; // h = 1
#pragma unroll
        for (int n = 0; n < NT; ++n)
        {
            d0[8 * n * EL + er]       = acc[n][0];
            d1[8 * n * EL + er]       = acc[n][1];
            d0[8 * n * EL + (er ^ 8)] = acc[n][2];
            d1[8 * n * EL + (er ^ 8)] = acc[n][3];
        }
    };

#pragma unroll 1
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
        __syncthreads(); // every warp has left direction 1 of the previous tile (and the fragments are built)
        issue(tile);
        cp_async_wait<0>();
        __syncthreads();

        uint4 f[KS][NT];
        load_frags(f, fr0);
        // direction 0: rows q = warp + k*WARPS, in place (a row belongs to one warp)
        {
            float a[2][KS][4], acc[2][NT][4];
            load_a(a[0], s1 + (warp * NQ + t) * EL, e0row, EL);
#pragma unroll
            for (int k = 0; k < C::ROWS; ++k)
            {
                const int q = warp + k * WARPS;
                if (k + 1 < C::ROWS && q + WARPS < NM)
                    load_a(a[(k + 1) & 1], s1 + ((q + WARPS) * NQ + t) * EL, e0row, EL);
                if (q < NM)
                    mma_row(acc[k & 1], a[k & 1], f);
                if (k > 0)
                    store_row(q - WARPS, acc[(k - 1) & 1]);
            }
            if (warp + (C::ROWS - 1) * WARPS < NM)
                store_row(warp + (C::ROWS - 1) * WARPS, acc[(C::ROWS - 1) & 1]);
        }
        __syncthreads();

        load_frags(f, fr1);
        // direction 1: columns i = warp + k*WARPS, straight to global
        {
            const unsigned group = tile / PER, l0 = (tile % PER) * EL;
            float *gout          = out + (size_t)group * 32 * C::NQ2 + l0 + g;
            // u = 32 (8 ks + t (+ 4)) + i: bit 1 = i >> 1, bit 2 = i >> 2, bit 5 = t & 1, bit 6 = t >> 1
            auto col_ptr = [&](int i) { return s1 + (t * NQ + (i ^ (((i >> 1) ^ t) & 1))) * EL; };
            auto col_e0  = [&](int i) { return g ^ (8 * (((i >> 2) ^ (t >> 1)) & 1)); };
            auto store_col = [&](int i, const float (&acc)[NT][4]) {
                float *dst = gout + (size_t)32 * ((2 * t) * NQ + i); // j = 8 n + 2 t + h
#pragma unroll
                for (int n = 0; n < NT; ++n)
                {
                    st_stream(dst + (size_t)32 * NQ * (8 * n), acc[n][0]);
                    st_stream(dst + (size_t)32 * NQ * (8 * n + 1), acc[n][1]);
                    st_stream(dst + (size_t)32 * NQ * (8 * n) + 8, acc[n][2]);
                    st_stream(dst + (size_t)32 * NQ * (8 * n + 1) + 8, acc[n][3]);
                }
            };
            float a[2][KS][4], acc[2][NT][4];
            load_a(a[0], col_ptr(warp), col_e0(warp), NQ * EL);
#pragma unroll
            for (int k = 0; k < C::COLS; ++k)
            {
                const int i = warp + k * WARPS;
                if (k + 1 < C::COLS && i + WARPS < NQ)
                    load_a(a[(k + 1) & 1], col_ptr(i + WARPS), col_e0(i + WARPS), NQ * EL);
                if (i < NQ)
                    mma_row(acc[k & 1], a[k & 1], f);
                if (k > 0)
                    store_col(i - WARPS, acc[(k - 1) & 1]);
            }
            if (warp + (C::COLS - 1) * WARPS < NQ)
                store_col(warp + (C::COLS - 1) * WARPS, acc[(C::COLS - 1) & 1]);
        }
    }
}

} // namespace b200fe

// sumfac_lanes.cuh -- "lanes" back-end for the warp-interleaved layout of the reference's
// "Coales" kernels (benchmark04.cc:78-147, benchmark05.cc:104-201)
//     x[(e/32)*32*len + 32*idx + e%32]
// at the nq where one thread can no longer hold a whole element (sumfac_tpe.cuh).
//
// The lanes of a warp still run over the elements of an interleave group, so every global access
// is a run of EL consecutive values of one in-element index (whole 128-byte lines for EL = 32, or
// for EL = 16 doubles) and every shared-memory access is conflict-free by construction -- no
// gather into an element-major tile as in sumfac_rows_coa.cuh.  What is split over the warps of a
// CTA is the element's work:
//   hex   phase A  worker r owns the plane in[r][.][.] (nm^2 registers, read from HBM with all the
//                  loads of the plane in flight), contracts directions 0 and 1 in registers and
//                  leaves t2[r][j][i] in shared memory;
//         phase B  worker w takes (j, i) rows w, w + NW, ...: nm values from shared memory,
//                  direction 2, nq coalesced stores.
//   quad  phase A  worker q owns the row in[q][.] -> t1[q][i] in shared memory;
//         phase B  worker i owns the column t1[.][i] -> out[.][i], stored straight to global.
// Summation order is the reference's (p, then q, then r, ascending, starting from +0), the basis
// operand comes from the constant bank, FP32 accumulators are FFMA2 pairs (RowAcc, sumfac_rows.cuh).
// Shared memory holds only the last intermediate: nm*nq^2*EL values (hex), nm*nq*EL (quad).
#pragma once

#include "sumfac_rows.cuh"

namespace b200fe
{

// output block of a row: as unrolled_ib (sumfac_rows.cuh), with the uniform-register budget stretched to 64 so that a
// row of 16 doubles (IProductWRTBase nq = 16) still gets 2-wide blocks
template <int NM, int NQ, int SIZE> constexpr int lanes_ib()
{
    for (int ib = 8; ib >= 2; ib /= 2)
        if (ib <= NQ && NM * ib * (SIZE / 4) <= 64)
            return ib;
    return 1;
}

// NQ outputs of one row of NM register values against bank matrix BOFF, in blocks of IB outputs;
// st(j, value) consumes output j
template <typename T, int NM, int NQ, int BOFF, typename St>
__device__ __forceinline__ void lanes_row(const T (&x)[NM], St st)
{
    constexpr int PITCH = bank_pitch<T>(NQ);
    if constexpr (lanes_ib<NM, NQ, (int)sizeof(T)>() < 2)
    {
        // very long rows (nq = 32): the basis of even a 2-wide block exceeds the uniform registers, so the
        // loop over the output blocks stays a loop (register-indexed 8-byte uniform loads, one per FMA pair)
        constexpr int RB = 8;
        static_assert(NQ % RB == 0, "rolled blocks");
#pragma unroll 1
        for (int ib = 0; ib < NQ; ib += RB)
        {
            RowAcc<T, RB> t;
            t.zero();
#pragma unroll
            for (int p = 0; p < NM; ++p)
            {
                T b[RB];
                cbasis_load<RB, (sizeof(T) == 4)>(BOFF + p * PITCH + ib, b);
                t.fma(x[p], b);
            }
#pragma unroll
            for (int j = 0; j < RB; ++j)
                st(ib + j, t.get(j));
        }
    }
    else
    {
        constexpr int IB = lanes_ib<NM, NQ, (int)sizeof(T)>();
#pragma unroll
        for (int ib = 0; ib + IB <= NQ; ib += IB)
        {
            RowAcc<T, IB> t;
            t.zero();
#pragma unroll
            for (int p = 0; p < NM; ++p)
            {
                T b[IB];
                cbasis_load<IB, true>(BOFF + p * PITCH + ib, b);
                t.fma(x[p], b);
            }
#pragma unroll
            for (int j = 0; j < IB; ++j)
                st(ib + j, t.get(j));
        }
        constexpr int TAIL = NQ % IB;
        if constexpr (TAIL > 0)
        {
            // FP32 with an odd number of outputs left (IProductWRTBase: nm outputs): the bank rows are padded with
            // zeros up to a whole 16-byte vector, so the tail runs one output wider on packed FMAs and drops the last
            constexpr int TW = (sizeof(T) == 4 && TAIL % 2 == 1 && NQ < PITCH) ? TAIL + 1 : TAIL;
            RowAcc<T, TW> t;
            t.zero();
#pragma unroll
            for (int p = 0; p < NM; ++p)
            {
                T b[TW];
                cbasis_load<TW, true>(BOFF + p * PITCH + (NQ - TAIL), b);
                t.fma(x[p], b);
            }
#pragma unroll
            for (int j = 0; j < TAIL; ++j)
                st(NQ - TAIL + j, t.get(j));
        }
    }
}

template <typename T, int NQ, int EL> struct QuadLanes
{
    static_assert(EL == 8 || EL == 16 || EL == 32, "a tile is a power-of-two slice of an interleave group");
    static constexpr int NM = NQ - 1, NM2 = NM * NM, NQ2 = NQ * NQ;
    static constexpr int THREADS = EL * NQ;
    static constexpr size_t SMEM = (size_t)NM * NQ * EL * sizeof(T);
    static_assert(THREADS <= 1024 && THREADS % 32 == 0, "block size");
};

template <typename T, int NQ, int EL>
__device__ __noinline__ void quad_lanes_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int EL>
__global__ void __launch_bounds__(QuadLanes<T, NQ, EL>::THREADS)
    bwdtrans_quad_lanes_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    quad_lanes_body<T, NQ, EL>(in, out, nelmt);
}

template <typename T, int NQ, int EL>
__device__ __noinline__ void quad_lanes_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    using C           = QuadLanes<T, NQ, EL>;
    constexpr int NM  = C::NM;
    constexpr int BP  = bank_pitch<T>(NQ), B0 = 0, B1 = NM * BP;
    constexpr int PER = 32 / EL; // tiles per interleave group
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s1 = reinterpret_cast<T *>(smem_raw);
    (void)nelmt; // nelmt % 32 == 0 is checked at the C ABI: every tile is full

    const int e = threadIdx.x % EL, w = threadIdx.x / EL;
    const unsigned group = blockIdx.x / PER, l0 = (blockIdx.x % PER) * EL;
    const T *gin = in + (size_t)group * 32 * C::NM2 + l0 + e;
    T *gout      = out + (size_t)group * 32 * C::NQ2 + l0 + e;

    if (w < NM)
    {
        T x[NM];
        const T *px = gin + (size_t)w * (32 * NM);
#pragma unroll
        for (int p = 0; p < NM; ++p)
            x[p] = ld_stream(px + 32 * p);
        T *dst = s1 + (size_t)(w * NQ) * EL + e;
        lanes_row<T, NM, NQ, B0>(x, [&](int i, T v) { dst[i * EL] = v; });
    }
    __syncthreads();
    {
        T x[NM];
        const T *src = s1 + (size_t)w * EL + e;
#pragma unroll
        for (int q = 0; q < NM; ++q)
            x[q] = src[q * NQ * EL];
        T *dst = gout + 32 * w;
        lanes_row<T, NM, NQ, B1>(x, [&](int j, T v) { st_stream(dst + 32 * NQ * j, v); });
    }
}

// directions 0 and 1 of one plane a[NIN][NIN] for the IBW outputs i in [ib, ib + IBW): t1[q][i] for every q in registers
// (the basis values of a (p, i-block) are shared by the NIN rows q), then column i of t1 against the second bank
// matrix -> t2[j][i] at dst[(j*NOUT + i)*S].  BwdTrans: NIN = nm, NOUT = nq; IProductWRTBase: NIN = nq, NOUT = nm
// with the transposed bank (common.cuh) -- the bank offsets are the same expressions in both.
template <typename T, int NIN, int NOUT, int S, int IBW>
__device__ __forceinline__ void plane_block(const T (&a)[NIN * NIN], T *dst, int ib)
{
    constexpr int BP = bank_pitch<T>(NOUT), B0 = 0, B1 = NIN * BP;
    RowAcc<T, IBW> t1[NIN];
#pragma unroll
    for (int q = 0; q < NIN; ++q)
        t1[q].zero();
#pragma unroll
    for (int p = 0; p < NIN; ++p)
    {
        T b[IBW];
        cbasis_load<IBW, true>(B0 + p * BP + ib, b);
#pragma unroll
        for (int q = 0; q < NIN; ++q)
            t1[q].fma(a[q * NIN + p], b);
    }
#pragma unroll
    for (int ii = 0; ii < IBW; ++ii)
    {
        T x[NIN];
#pragma unroll
        for (int q = 0; q < NIN; ++q)
            x[q] = t1[q].get(ii);
        T *d = dst + (ib + ii) * S;
        lanes_row<T, NIN, NOUT, B1>(x, [&](int j, T v) { d[j * NOUT * S] = v; });
    }
}
template <typename T, int NQ, int EL, int IBW>
__device__ __forceinline__ void hex_lanes_block(const T (&a)[(NQ - 1) * (NQ - 1)], T *dst, int ib)
{
    plane_block<T, NQ - 1, NQ, EL, IBW>(a, dst, ib);
}

template <typename T, int NQ, int EL> struct HexLanes
{
    static_assert(EL == 8 || EL == 16 || EL == 32, "a tile is a power-of-two slice of an interleave group");
    static constexpr int NM = NQ - 1, NM2 = NM * NM, NM3 = NM2 * NM, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
    static constexpr int THREADS = (EL * NM + 31) / 32 * 32;
    static constexpr int NW      = THREADS / EL; // workers per element (>= NM)
    static constexpr size_t SMEM = (size_t)NM * NQ2 * EL * sizeof(T);
    // direction-0 outputs kept per register block: pairs for the packed FP32 FMA
    static constexpr int IB0 = sizeof(T) == 4 ? 2 : 1;
    // FP64: the loop over the direction-0 output blocks stays a loop.  Unrolled, ptxas computes t1 for every i
    // at once (nm*nq live doubles on top of the plane: 230 registers at nq = 8, spills at nq = 10).
    static constexpr bool ROLLED = sizeof(T) == 8 || NQ >= 8; // (FP32: 122 registers unrolled at nq = 8, 182-255 beyond)
};

template <typename T, int NQ, int EL, int MINB>
__device__ __noinline__ void hex_lanes_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int EL, int MINB = 1>
__global__ void __launch_bounds__(HexLanes<T, NQ, EL>::THREADS, MINB)
    bwdtrans_hex_lanes_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    hex_lanes_body<T, NQ, EL, MINB>(in, out, nelmt);
}

template <typename T, int NQ, int EL, int MINB>
__device__ __noinline__ void hex_lanes_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    using C           = HexLanes<T, NQ, EL>;
    constexpr int NM = C::NM, NM2 = C::NM2, NQ2 = C::NQ2, NW = C::NW, IB0 = C::IB0;
    constexpr int BP  = bank_pitch<T>(NQ), B2 = 2 * NM * BP;
    constexpr int PER = 32 / EL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s2 = reinterpret_cast<T *>(smem_raw);
    (void)nelmt;

    const int e = threadIdx.x % EL, w = threadIdx.x / EL;
    const unsigned group = blockIdx.x / PER, l0 = (blockIdx.x % PER) * EL;
    const T *gin = in + (size_t)group * 32 * C::NM3 + l0 + e;
    T *gout      = out + (size_t)group * 32 * C::NQ3 + l0 + e; // intended group stride (benchmark05.cc:810-812)

    if (w < NM)
    {
        // plane r = w of the element, whole in registers
        T a[NM2];
        const T *pa = gin + (size_t)w * (32 * NM2); // one base register, immediate offsets
#pragma unroll
        for (int k = 0; k < NM2; ++k)
            a[k] = ld_stream(pa + 32 * k);
        T *dst = s2 + (size_t)(w * NQ2) * EL + e;
        if constexpr (C::ROLLED)
        {
#pragma unroll 1
            for (int ib = 0; ib + IB0 <= NQ; ib += IB0)
                hex_lanes_block<T, NQ, EL, IB0>(a, dst, ib);
        }
        else
        {
#pragma unroll
            for (int ib = 0; ib + IB0 <= NQ; ib += IB0)
                hex_lanes_block<T, NQ, EL, IB0>(a, dst, ib);
        }
        if constexpr (NQ % IB0 != 0) // odd nq in FP32: the last output alone
            hex_lanes_block<T, NQ, EL, 1>(a, dst, NQ - 1);
    }
    __syncthreads();

    // direction 2: rows (j, i) of t2 over r, nq outputs each straight to global
    constexpr int ITER = (NQ2 + NW - 1) / NW;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it)
    {
        const int ji = w + it * NW;
        if (ji >= NQ2)
            break;
        T x[NM];
        const T *src = s2 + (size_t)ji * EL + e;
#pragma unroll
        for (int r = 0; r < NM; ++r)
            x[r] = src[r * NQ2 * EL];
        T *dst = gout + 32 * ji;
        lanes_row<T, NM, NQ, B2>(x, [&](int k, T v) { st_stream(dst + 32 * NQ2 * k, v); });
    }
}

// ---- hex, q-outer variant for the nq whose plane no longer fits the register file (nq = 10): worker (r, ih) streams
// the rows in[r][q][.] of its plane one at a time and keeps the accumulators t2[r][.][i] of its slice of NQ/IH outputs i
// (fma over q ascending: the reference's order), so only nq*nq/IH accumulators + one row are live.  The IH workers of
// a plane read the same rows (the second read hits L1).
template <typename T, int NQ, int EL, int IH, int I0>
__device__ __forceinline__ void hex_lanes_plane_q(const T *pa, T *dst)
{
    constexpr int NM = NQ - 1, NI = NQ / IH, W = 16 / (int)sizeof(T);
    constexpr int BP = bank_pitch<T>(NQ), B0 = 0, B1 = NM * BP;
    RowAcc<T, NQ> acc[NI];
#pragma unroll
    for (int ii = 0; ii < NI; ++ii)
        acc[ii].zero();
#pragma unroll
    for (int q = 0; q < NM; ++q)
    {
        T a[NM];
#pragma unroll
        for (int p = 0; p < NM; ++p)
            a[p] = ld_stream(pa + 32 * (q * NM + p));
        RowAcc<T, NI> t;
        t.zero();
#pragma unroll
        for (int p = 0; p < NM; ++p)
        {
            T b[NI];
            cbasis_load<NI, (I0 % W == 0)>(B0 + p * BP + I0, b);
            t.fma(a[p], b);
        }
        T b1[NQ];
        cbasis_load<NQ, true>(B1 + q * BP, b1);
#pragma unroll
        for (int ii = 0; ii < NI; ++ii)
            acc[ii].fma(t.get(ii), b1);
    }
#pragma unroll
    for (int ii = 0; ii < NI; ++ii)
#pragma unroll
        for (int j = 0; j < NQ; ++j)
            dst[(j * NQ + I0 + ii) * EL] = acc[ii].get(j);
}

template <typename T, int NQ, int EL, int IH> struct HexLanesQ
{
    static constexpr int NM = NQ - 1, NM2 = NM * NM, NM3 = NM2 * NM, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
    static constexpr int WPI     = (NM * EL + 31) / 32 * 32 / EL; // workers per i-slice: whole warps, so a warp's slice is uniform
    static constexpr int NW      = WPI * IH;
    static constexpr int THREADS = NW * EL;
    static constexpr size_t SMEM = (size_t)NM * NQ2 * EL * sizeof(T);
    static_assert(NQ % IH == 0, "");
};

template <typename T, int NQ, int EL, int IH, int MINB>
__device__ __noinline__ void hex_lanesq_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt);

template <typename T, int NQ, int EL, int IH, int MINB>
__global__ void __launch_bounds__(HexLanesQ<T, NQ, EL, IH>::THREADS, MINB)
    bwdtrans_hex_lanesq_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    hex_lanesq_body<T, NQ, EL, IH, MINB>(in, out, nelmt);
}

template <typename T, int NQ, int EL, int IH, int MINB>
__device__ __noinline__ void hex_lanesq_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt)
{
    using C           = HexLanesQ<T, NQ, EL, IH>;
    constexpr int NM = C::NM, NM2 = C::NM2, NQ2 = C::NQ2, NW = C::NW, NI = NQ / IH;
    constexpr int BP  = bank_pitch<T>(NQ), B2 = 2 * NM * BP;
    constexpr int PER = 32 / EL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s2 = reinterpret_cast<T *>(smem_raw);
    (void)nelmt;
    const int e = threadIdx.x % EL, w = threadIdx.x / EL;
    const unsigned group = blockIdx.x / PER, l0 = (blockIdx.x % PER) * EL;
    const T *gin = in + (size_t)group * 32 * C::NM3 + l0 + e;
    T *gout      = out + (size_t)group * 32 * C::NQ3 + l0 + e;
    const int ih = w / C::WPI, r = w - ih * C::WPI;
    if (r < NM)
    {
        const T *pa = gin + (size_t)r * (32 * NM2);
        T *dst      = s2 + (size_t)(r * NQ2) * EL + e;
        if constexpr (IH == 1)
            hex_lanes_plane_q<T, NQ, EL, IH, 0>(pa, dst);
        else if constexpr (IH == 2)
        {
            if (ih == 0)
                hex_lanes_plane_q<T, NQ, EL, IH, 0>(pa, dst);
            else
                hex_lanes_plane_q<T, NQ, EL, IH, NI>(pa, dst);
        }
    }
    __syncthreads();
    constexpr int ITER = (NQ2 + NW - 1) / NW;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it)
    {
        const int ji = w + it * NW;
        if (ji >= NQ2)
            break;
        T x[NM];
        const T *src = s2 + (size_t)ji * EL + e;
#pragma unroll
        for (int rr = 0; rr < NM; ++rr)
            x[rr] = src[rr * NQ2 * EL];
        T *dst = gout + 32 * ji;
        lanes_row<T, NM, NQ, B2>(x, [&](int k, T v) { st_stream(dst + 32 * NQ2 * k, v); });
    }
}


// ---- element-major quads in the same style ("lanes-em") -------------------------------------------------------
// The CTA's slab of EL elements is contiguous in the element-major layout: one bulk (TMA) copy brings it into shared
// memory unchanged, and because nm^2 is odd for even nq the lanes of a warp (= elements, stride nm^2) read it without
// bank conflicts.  Direction 0 is the interleaved kernel's phase A (worker q owns row q of its lane's element);
// t1 goes to shared memory as [q][e][i] with the element stride padded to an odd nq + 1 (conflict-free stores).
// Direction 1 re-maps the threads to (e, i) flattened, i fastest -- the basis operand depends on neither -- so that
// the nq outputs of a thread go straight to global memory as runs of nq consecutive values per element.
// One row per thread and phase, no loops, no index arithmetic beyond the prologue: the FP32 cases that the
// general row kernel (sumfac_rows.cuh) leaves issue-bound (nq = 14: 0.66 of the roofline) are the target.
// TPC > 1: the CTA takes TPC consecutive tiles.  All their bulk copies are issued up front (TPC slots, one mbarrier
// each), so the later tiles' loads are in flight while the first is contracted -- for the small nq, where a tile is
// only 6-14 KB and a one-tile CTA is latency-bound.  t1 is double-buffered: one barrier per tile.
template <typename T, int NQ, int EL, int TPC = 1> struct QuadLanesEm
{
    static_assert(NQ % 2 == 0, "nm^2 must be odd for conflict-free lane access to the unpadded slab");
    static constexpr int NM = NQ - 1, NM2 = NM * NM, NQ2 = NQ * NQ;
    static constexpr int WORK    = EL * NQ;                // threads with a row to contract
    static constexpr int THREADS = (WORK + 31) / 32 * 32;  // whole warps: cta_sum_fixed shuffles over full warps
    static constexpr int E1      = NQ + 1;  // element stride of t1 (odd)
    static constexpr int Q1      = EL * E1; // row stride of t1
    static constexpr int SIN     = (EL * NM2 * (int)sizeof(T) + 15) / 16 * 16 / (int)sizeof(T);
    static constexpr int S1      = (NM * Q1 * (int)sizeof(T) + 15) / 16 * 16 / (int)sizeof(T);
    static constexpr int NBUF    = TPC > 1 ? 2 : 1;
    static constexpr size_t SMEM = (size_t)(TPC * SIN + NBUF * S1) * sizeof(T) + 8 * ((TPC + 1) / 2 * 2);
    static_assert(THREADS <= 1024 && THREADS % 32 == 0, "block size: whole warps");
};

// SUMSQ: operator + checksum fused: every thread squares what it stores, the CTA leaves one partial
// (fixed summation order) in partials[blockIdx.x]
template <typename T, int NQ, int EL, int MINB, int TPC, bool SUMSQ>
__device__ __noinline__ void quad_lanesem_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt,
                                 double *__restrict__ partials);

template <typename T, int NQ, int EL, int MINB = 1, int TPC = 1, bool SUMSQ = false>
__global__ void __launch_bounds__(QuadLanesEm<T, NQ, EL, TPC>::THREADS, MINB)
    bwdtrans_quad_lanesem_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt,
                                 double *__restrict__ partials)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    quad_lanesem_body<T, NQ, EL, MINB, TPC, SUMSQ>(in, out, nelmt, partials);
}

template <typename T, int NQ, int EL, int MINB, int TPC, bool SUMSQ>
__device__ __noinline__ void quad_lanesem_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt,
                                 double *__restrict__ partials)
{
    double ss = 0.0;
    using C = QuadLanesEm<T, NQ, EL, TPC>;
    constexpr int NM = C::NM, NM2 = C::NM2, NQ2 = C::NQ2, E1 = C::E1, Q1 = C::Q1;
    constexpr int BP = bank_pitch<T>(NQ), B0 = 0, B1 = NM * BP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_in       = reinterpret_cast<T *>(smem_raw);
    T *s1         = s_in + TPC * C::SIN;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)(TPC * C::SIN + C::NBUF * C::S1) * sizeof(T));

    const int tid        = threadIdx.x;
    const unsigned tile0 = blockIdx.x * TPC;
    if (tid == 0)
    {
#pragma unroll
        for (int t = 0; t < TPC; ++t)
            mbar_init(bar + t, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
    {
#pragma unroll
        for (int t = 0; t < TPC; ++t)
            if ((size_t)(tile0 + t) * EL < nelmt)
                ring_issue<T, EL, NM2>(s_in + t * C::SIN, bar + t, in, tile0 + t, nelmt);
    }
#pragma unroll
    for (int t = 0; t < TPC; ++t)
    {
        const size_t e0 = (size_t)(tile0 + t) * EL;
        if (e0 >= nelmt) // uniform over the CTA
            break;
        const int ne = (nelmt - e0 < (size_t)EL) ? (int)(nelmt - e0) : EL;
        const T *sl  = s_in + t * C::SIN;
        T *st1       = s1 + (t & 1) * C::S1;
        ring_wait<T, NM2>(s_in + t * C::SIN, bar + t, 0u, in + e0 * NM2, ne, tid);
        {
            const int e = tid % EL, q = tid / EL;
            if (q < NM)
            {
                T x[NM];
                const T *src = sl + e * NM2 + q * NM;
#pragma unroll
                for (int p = 0; p < NM; ++p)
                    x[p] = src[p];
                T *dst = st1 + q * Q1 + e * E1;
                lanes_row<T, NM, NQ, B0>(x, [&](int i, T v) { dst[i] = v; });
            }
        }
        __syncthreads();
        {
            const int e = tid / NQ, i = tid - e * NQ;
            T x[NM];
            const T *src = st1 + e * E1 + i;
            if (tid < C::WORK) // the threads that pad the last warp own no row: they only take part in the barriers
            {
#pragma unroll
                for (int q = 0; q < NM; ++q)
                    x[q] = src[q * Q1];
            }
            if (tid < C::WORK && e < ne)
            {
                T *dst = out + (e0 + e) * NQ2 + i;
                lanes_row<T, NM, NQ, B1>(x, [&](int j, T v) {
                    st_stream(dst + j * NQ, v);
                    if (SUMSQ)
                        ss = fma((double)v, (double)v, ss);
                });
            }
        }
    }
    if constexpr (SUMSQ)
    {
        __shared__ double red[32];
        const double total = cta_sum_fixed(ss, red);
        if (tid == 0)
            partials[blockIdx.x] = total;
    }
}

// ---- element-major hexes in the same style ("lanes-em") -------------------------------------------------------
// As for the quads: the CTA's slab of EL elements arrives with one bulk copy and is read with lanes = elements
// (stride nm^3, odd for even nq: conflict-free).  Worker r takes its plane into registers; after a barrier the slab's
// shared memory is reused for t2[e][r][(j, i)] (element stride padded to an odd count, so the lanes = elements stores
// of phase A and the lanes = (j, i) loads of phase B are both conflict-free).  Phase B flattens (e, (j, i)) over the
// threads, (j, i) fastest: for every k a warp stores 32 consecutive values of out[e][k][.][.].
template <typename T, int NQ, int EL> struct HexLanesEm
{
    static_assert(NQ % 2 == 0, "nm^3 must be odd for conflict-free lane access to the unpadded slab");
    static constexpr int NM = NQ - 1, NM2 = NM * NM, NM3 = NM2 * NM, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
    static constexpr int THREADS = (EL * NM + 31) / 32 * 32;
    static constexpr int ES      = NM * NQ2 + 1; // element stride of t2: nm * nq^2 is even
    static constexpr int SIN     = (EL * NM3 * (int)sizeof(T) + 15) / 16 * 16 / (int)sizeof(T);
    static constexpr int SBUF    = SIN > EL * ES ? SIN : EL * ES;
    static constexpr size_t SMEM = (size_t)SBUF * sizeof(T) + 16;
    static constexpr int IB0     = sizeof(T) == 4 ? 2 : 1;
    static constexpr bool ROLLED = sizeof(T) == 8 || NQ >= 8; // see HexLanes
};

template <typename T, int NQ, int EL, int MINB, bool SUMSQ>
__device__ __noinline__ void hex_lanesem_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt,
                                double *__restrict__ partials);

template <typename T, int NQ, int EL, int MINB = 1, bool SUMSQ = false>
__global__ void __launch_bounds__(HexLanesEm<T, NQ, EL>::THREADS, MINB)
    bwdtrans_hex_lanesem_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt,
                                double *__restrict__ partials)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    hex_lanesem_body<T, NQ, EL, MINB, SUMSQ>(in, out, nelmt, partials);
}

template <typename T, int NQ, int EL, int MINB, bool SUMSQ>
__device__ __noinline__ void hex_lanesem_body(const T *__restrict__ in, T *__restrict__ out, unsigned nelmt,
                                double *__restrict__ partials)
{
    double ss = 0.0;
    using C = HexLanesEm<T, NQ, EL>;
    constexpr int NM = C::NM, NM2 = C::NM2, NM3 = C::NM3, NQ2 = C::NQ2, ES = C::ES, IB0 = C::IB0;
    constexpr int BP = bank_pitch<T>(NQ), B2 = 2 * NM * BP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_in       = reinterpret_cast<T *>(smem_raw);
    T *s2         = s_in; // reused once every worker holds its plane
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)C::SBUF * sizeof(T));

    const int tid   = threadIdx.x;
    const size_t e0 = (size_t)blockIdx.x * EL;
    const int ne    = (nelmt - e0 < (size_t)EL) ? (int)(nelmt - e0) : EL;
    if (tid == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
        ring_issue<T, EL, NM3>(s_in, bar, in, blockIdx.x, nelmt);
    ring_wait<T, NM3>(s_in, bar, 0u, in + e0 * NM3, ne, tid);

    const int e = tid % EL, r = tid / EL;
    T a[NM2];
    if (r < NM)
    {
        const T *src = s_in + e * NM3 + r * NM2;
#pragma unroll
        for (int k = 0; k < NM2; ++k)
            a[k] = src[k];
    }
    __syncthreads();
    if (r < NM)
    {
        T *dst = s2 + e * ES + r * NQ2;
        if constexpr (C::ROLLED)
        {
#pragma unroll 1
            for (int ib = 0; ib + IB0 <= NQ; ib += IB0)
                hex_lanes_block<T, NQ, 1, IB0>(a, dst, ib);
        }
        else
        {
#pragma unroll
            for (int ib = 0; ib + IB0 <= NQ; ib += IB0)
                hex_lanes_block<T, NQ, 1, IB0>(a, dst, ib);
        }
    }
    __syncthreads();

    constexpr int ROWS = EL * NQ2, ITER = (ROWS + C::THREADS - 1) / C::THREADS;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it)
    {
        const int row = it * C::THREADS + tid;
        if (row >= ROWS)
            break;
        const int e2 = row / NQ2, ji = row - e2 * NQ2;
        T x[NM];
        const T *src = s2 + e2 * ES + ji;
#pragma unroll
        for (int rr = 0; rr < NM; ++rr)
            x[rr] = src[rr * NQ2];
        if (e2 < ne)
        {
            T *dst = out + (e0 + e2) * C::NQ3 + ji;
            lanes_row<T, NM, NQ, B2>(x, [&](int k, T v) {
                st_stream(dst + k * NQ2, v);
                if (SUMSQ)
                    ss = fma((double)v, (double)v, ss);
            });
        }
    }
    if constexpr (SUMSQ)
    {
        __shared__ double red[32];
        const double total = cta_sum_fixed(ss, red);
        if (tid == 0)
            partials[blockIdx.x] = total;
    }
}

} // namespace b200fe

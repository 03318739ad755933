// sumfac_coapipe.cuh -- "coa-pipe" back-end: hexes in the warp-interleaved layout of the reference's "Coales"
// kernel (benchmark05.cc:104-201, offsets :810-812)
//     x[(e/32)*32*len + 32*idx + e%32]
// at the nq where a thread can no longer keep a plane of its element in registers (nq = 10: 81 values; the
// q-outer lanes kernel that served it is latency-bound at 0.46 of the roofline -- profiles/r02_ncu_hex10_f64_coa.txt).
//
// The tile of a CTA is EL consecutive elements of one interleave group and it keeps the group's index order
// [idx][e] ALL THE WAY: in shared memory, through both intermediates and out to global memory, so nothing is
// ever transposed.  The lanes of a warp run over the elements (and 32/EL consecutive rows), a thread contracts one
// row at a time (nm values in, nq out -- lanes_row of sumfac_lanes.cuh, basis operand from the constant bank,
// warp-uniform because every lane works on the same kind of row):
//     pass 0   rows (r, q):  in[r][q][.]  -> s1[r][q][i]
//     pass 1   rows (r, i):  s1[r][.][i]  -> s2[r][j][i]
//     pass 2   rows (j, i):  s2[.][j][i]  -> out[.][j][i]   straight to global, runs of EL values
// The gather of the tile (runs of EL*sizeof(T) = 64 / 128 bytes, 256 bytes apart) is TILED TMA: the interleaved array is
// described to the copy engine as a rank-3 tensor {32 elements, nm^3 indices, groups} (tensormap.h) and a box of
// {EL, 256, 1} lands dense, as [idx][e], in the slot -- one cp.async.bulk.tensor instruction per 256 indices, issued by
// one thread, completion on an mbarrier (SASS UTMALDG).  Without the driver entry point that encodes the map the same
// kernel gathers with 16-byte cp.async copies issued by every thread (0.70 instead of 0.74 at nq = 10: the copies compete
// with the passes for the LSU / MIO queues).  Persistent CTAs; shared memory is the input slot
// plus ONE work region of nm planes x nq x nq indices: pass 0 goes slot -> work (s1[r][q][i], nm of a plane's nq
// rows), after which the slot is free and the next tile's gather is issued at once (it lands under passes 1 and 2);
// pass 1 runs IN PLACE, column by column (s2[r][j][i] overwrites s1[r][q = j][i], a thread owns its column, no barrier
// inside the pass); pass 2 goes work -> global.  Three barriers per tile.  107 KB at nq = 10 FP64 with 8 elements
// per tile: two CTAs per SM, each with its own prefetch.
// Bank conflicts: a warp touches 32/EL rows at once, U = EL*sizeof(T) bytes each (64 here: two per bank window).  The
// row stride of the slot is nm*U (nm odd: consecutive rows alternate between the halves of a window).  The work
// region's row stride is nq*U (even nq: every row on the same half), so index i of row rho = r*nm + q is stored at
// i ^ (rho & 1): the stores of pass 0 (consecutive rho, one i), the loads and stores of pass 1 (one rho, consecutive
// i) and the loads of pass 2 are all conflict-free.
// Summation order is the reference's (p, then q, then r, ascending from +0, fused multiply-adds): bit-identical.
#pragma once

#include "sumfac_lanes.cuh"
#include "tensormap.h"

namespace b200fe
{

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// one box of a rank-3 tensor map -> shared memory (tiled TMA), completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_addr(smem_dst)),
                 "l"(map), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// predicated stores: a store behind `if (ok)` is a branch, and with a branch after every row's block ptxas contracts
// the R rows of a pass one after the other (constants reloaded per row, two dependent DFMA chains in flight); a
// predicated store keeps the pass one basic block, so the rows really share the loads and interleave their chains
__device__ __forceinline__ void st_shared_if(double *p, double v, bool ok)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.f64 [%0], %1; }" ::"r"(smem_addr(p)), "d"(v), "r"((unsigned)ok)
                 : "memory");
}
__device__ __forceinline__ void st_shared_if(float *p, float v, bool ok)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.f32 [%0], %1; }" ::"r"(smem_addr(p)), "f"(v), "r"((unsigned)ok)
                 : "memory");
}
__device__ __forceinline__ void st_stream_if(double *p, double v, bool ok)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.global.cs.f64 [%0], %1; }" ::"l"(p), "d"(v), "r"((unsigned)ok) : "memory");
}
__device__ __forceinline__ void st_stream_if(float *p, float v, bool ok)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.global.cs.f32 [%0], %1; }" ::"l"(p), "f"(v), "r"((unsigned)ok) : "memory");
}

// R rows at once: NQ outputs of each of R rows of NM register values against bank matrix BOFF, in blocks of IB outputs
// whose basis values (one 16-byte uniform load per 2 doubles / 4 floats) are shared by the R rows; st(k, i, v) consumes
// output i of row k.  A pass is ONE straight-line block per tile, R = ceil(rows / workers): with a loop over the rows
// ptxas hoists the loop-invariant constant-bank loads out of it and spills them through vector registers (650 other
// instructions per 270 DFMAs at nq = 10), and R rows per load also take the uniform-load rate out of the FP64 pipe's
// way (profiles/r01_ubench_fp_pipe.txt: 27 TFLOP/s with one load per 2 DFMAs, 36.8 from registers).
template <typename T, int NM, int NQ, int BOFF, int R, typename St>
__device__ __forceinline__ void coa_rows(const T (&x)[R][NM], St st)
{
    constexpr int PITCH = bank_pitch<T>(NQ);
    constexpr int IB    = lanes_ib<NM, NQ, (int)sizeof(T)>();
    static_assert(IB >= 2, "row too long for the unrolled form");
#pragma unroll
    for (int ib = 0; ib + IB <= NQ; ib += IB)
    {
        RowAcc<T, IB> t[R];
#pragma unroll
        for (int k = 0; k < R; ++k)
            t[k].zero();
#pragma unroll
        for (int p = 0; p < NM; ++p)
        {
            T b[IB];
            cbasis_load<IB, true>(BOFF + p * PITCH + ib, b);
#pragma unroll
            for (int k = 0; k < R; ++k)
                t[k].fma(x[k][p], b);
        }
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int j = 0; j < IB; ++j)
                st(k, ib + j, t[k].get(j));
    }
    constexpr int TAIL = NQ % IB;
    if constexpr (TAIL > 0)
    {
        static_assert(TAIL % 2 == 0, "even nq");
        RowAcc<T, TAIL> t[R];
#pragma unroll
        for (int k = 0; k < R; ++k)
            t[k].zero();
#pragma unroll
        for (int p = 0; p < NM; ++p)
        {
            T b[TAIL];
            cbasis_load<TAIL, true>(BOFF + p * PITCH + (NQ - TAIL), b);
#pragma unroll
            for (int k = 0; k < R; ++k)
                t[k].fma(x[k][p], b);
        }
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int j = 0; j < TAIL; ++j)
                st(k, NQ - TAIL + j, t[k].get(j));
    }
}

template <typename T, int NQ, int EL, int NW> struct HexCoaPipe
{
    static_assert(NQ % 2 == 0, "the swizzle of s1 assumes an even nq (odd nq is conflict-free unpadded)");
    static_assert(EL * sizeof(T) == 32 || EL * sizeof(T) == 64 || EL * sizeof(T) == 128, "runs of whole sectors");
    static constexpr int NM = NQ - 1, NM2 = NM * NM, NM3 = NM2 * NM, NQ2 = NQ * NQ, NQ3 = NQ2 * NQ;
    static constexpr int U   = EL * (int)sizeof(T);  // bytes of one index of the tile
    static constexpr int G   = 128 / U;              // indices per 128-byte bank window
    static_assert(G <= 2, "the one-bit swizzle serves 64- and 128-byte runs");
    static constexpr int THREADS = (EL * NW + 31) / 32 * 32;
    static constexpr int SIN = NM3 * EL;       // the slot: the tile as it lies in memory
    static constexpr int SW  = NM * NQ2 * EL;  // the work region: planes of nq x nq indices (s1 fills nm of the nq rows)
    static constexpr size_t SMEM = (size_t)(SIN + SW) * sizeof(T);
    static constexpr int PER = 32 / EL; // tiles per interleave group
    // tiled-TMA gather: boxes of {EL, 256, 1} (a shared-memory destination must be 128-byte aligned, so the box height is
    // a power of two; the rows of the last box beyond nm^3 are out of bounds in the map and arrive as zeros)
    static constexpr int BOXR = 256, NBOX = (NM3 + BOXR - 1) / BOXR;
    static constexpr int SIN_TMA = NBOX * BOXR * EL;
    static constexpr size_t SMEM_TMA = (size_t)(SIN_TMA + SW) * sizeof(T) + 16; // + the mbarrier
};

// TMAP: the gather is NBOX tiled-TMA copies (cp.async.bulk.tensor through a tensor map of the interleaved array, issued
// by one thread, completion on an mbarrier) instead of 16-byte cp.async copies issued by every thread.
template <typename T, int NQ, int EL, int NW, int MINB, bool TMAP>
__device__ __noinline__ void hex_coapipe_body(const T *__restrict__ in, T *__restrict__ out, unsigned ntiles,
                                              const CUtensorMap *map);

template <typename T, int NQ, int EL, int NW, int MINB = 1>
__global__ void __launch_bounds__(HexCoaPipe<T, NQ, EL, NW>::THREADS, MINB)
    bwdtrans_hex_coapipe_kernel(const T *__restrict__ in, T *__restrict__ out, unsigned ntiles)
{
    pdl_wait(); // programmatic dependent of the bank fill; the body is a real call (common.cuh)
    hex_coapipe_body<T, NQ, EL, NW, MINB, false>(in, out, ntiles, nullptr);
}
template <typename T, int NQ, int EL, int NW, int MINB = 1>
__global__ void __launch_bounds__(HexCoaPipe<T, NQ, EL, NW>::THREADS, MINB)
    bwdtrans_hex_coapipe_tma_kernel(const __grid_constant__ CUtensorMap map, T *__restrict__ out, unsigned ntiles)
{
    pdl_wait();
    hex_coapipe_body<T, NQ, EL, NW, MINB, true>(nullptr, out, ntiles, &map);
}

template <typename T, int NQ, int EL, int NW, int MINB, bool TMAP>
__device__ __noinline__ void hex_coapipe_body(const T *__restrict__ in, T *__restrict__ out, unsigned ntiles,
                                              const CUtensorMap *map)
{
    using C           = HexCoaPipe<T, NQ, EL, NW>;
    constexpr int NM = C::NM, NM2 = C::NM2, NM3 = C::NM3, NQ2 = C::NQ2, G = C::G, PER = C::PER;
    constexpr int BP = bank_pitch<T>(NQ), B0 = 0, B1 = NM * BP, B2 = 2 * NM * BP;
    constexpr int VW = 16 / (int)sizeof(T), CH = C::U / 16; // values per 16-byte chunk, chunks per index
    extern __shared__ __align__(128) unsigned char smem_raw128[]; // (own name: the TU also declares smem_raw with 16-byte alignment)
    T *slot = reinterpret_cast<T *>(smem_raw128);
    T *wk   = slot + (TMAP ? C::SIN_TMA : C::SIN); // the work region: s1[r][q][i] after pass 0, s2[r][j][i] (in place) after pass 1

    const int tid = threadIdx.x, e = tid % EL;
    const int w   = tid / EL < NW ? tid / EL : (1 << 20); // the threads that pad the last warp own no row (ok = false)

    uint64_t *bar = reinterpret_cast<uint64_t *>(wk + C::SW); // (TMAP only)
    if constexpr (TMAP)
    {
        if (tid == 0)
        {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    auto issue = [&](unsigned tile) {
        const unsigned group = tile / PER, l0 = (tile % PER) * EL;
        if constexpr (TMAP)
        {
            if (tid == 0) // (called behind a barrier: every generic-proxy read of the slot is done)
            {
                fence_proxy_async();
                mbar_arrive_expect_tx(bar, (unsigned)(C::SIN_TMA * sizeof(T))); // whole boxes, the zero-filled rows too
#pragma unroll
                for (int b = 0; b < C::NBOX; ++b)
                    tma_load_3d(slot + b * C::BOXR * EL, map, (int)l0, b * C::BOXR, (int)group, bar);
            }
        }
        else
        {
            const T *g = in + (size_t)group * 32 * NM3 + l0;
#pragma unroll 4
            for (int c = tid; c < NM3 * CH; c += C::THREADS)
            {
                const int idx = c / CH, part = c - idx * CH;
                cp_async16(slot + c * VW, g + (size_t)idx * 32 + part * VW);
            }
            cp_async_commit();
        }
    };

    unsigned tile = blockIdx.x, it = 0;
    if (tile < ntiles)
        issue(tile);
#pragma unroll 1
    for (; tile < ntiles; tile += gridDim.x, ++it)
    {
        if constexpr (TMAP)
            mbar_wait(bar, it & 1u);
        else
            cp_async_wait<0>();
        __syncthreads(); // the tile has landed; every warp has left pass 2 of the previous tile (the work region is free)

        // pass 0: rows rho = (r, q); worker w takes rho = w, w + NW, ... all at once.  Row (r, q) of the work region
        // starts at index r*nq^2 + q*nq and holds output i at i ^ (rho & (G-1)).
        {
            constexpr int R = (NM2 + NW - 1) / NW;
            T x[R][NM];
            T *db[R][G];
            bool ok[R];
#pragma unroll
            for (int k = 0; k < R; ++k)
            {
                ok[k]         = w + k * NW < NM2;
                const int rho = ok[k] ? w + k * NW : 0; // (clamped: computes something harmless, stores nothing)
                const int r = rho / NM, q = rho - r * NM;
                const T *src  = slot + (rho * NM) * EL + e;
#pragma unroll
                for (int p = 0; p < NM; ++p)
                    x[k][p] = src[p * EL];
#pragma unroll
                for (int m = 0; m < G; ++m)
                    db[k][m] = wk + (r * NQ2 + q * NQ + (m ^ (rho & (G - 1)))) * EL + e;
            }
            coa_rows<T, NM, NQ, B0, R>(x, [&](int k, int i, T v) { st_shared_if(db[k][i & (G - 1)] + (i & ~(G - 1)) * EL, v, ok[k]); });
        }
        __syncthreads();
        // the slot is drained: the next tile's gather runs under passes 1 and 2
        if (tile + gridDim.x < ntiles)
            issue(tile + gridDim.x);

        // pass 1: columns (r, i) over q, IN PLACE and without a barrier: output j of column (r, i) goes where input
        // q = j came from (row j of the plane, index i ^ ((r*nm + j) & (G-1))), j = nm into the plane's spare row --
        // a thread only ever overwrites what it has read itself.
        {
            constexpr int R = (NM * NQ + NW - 1) / NW;
            T x[R][NM];
            T *cb[R][G];
            bool ok[R];
#pragma unroll
            for (int k = 0; k < R; ++k)
            {
                ok[k]        = w + k * NW < NM * NQ;
                const int ri = ok[k] ? w + k * NW : 0;
                const int r = ri / NQ, i = ri - r * NQ, rho0 = r * NM;
#pragma unroll
                for (int m = 0; m < G; ++m)
                    cb[k][m] = wk + (r * NQ2 + (i ^ ((rho0 + m) & (G - 1)))) * EL + e; // rows q with q % G == m
#pragma unroll
                for (int q = 0; q < NM; ++q)
                    x[k][q] = cb[k][q & (G - 1)][(q * NQ) * EL];
            }
            coa_rows<T, NM, NQ, B1, R>(x, [&](int k, int j, T v) { st_shared_if(cb[k][j & (G - 1)] + (j * NQ) * EL, v, ok[k]); });
        }
        __syncthreads();

        // pass 2: rows (j, i) over r, straight to global
        {
            constexpr int R      = (NQ2 + NW - 1) / NW;
            const unsigned group = tile / PER, l0 = (tile % PER) * EL;
            T *gout              = out + (size_t)group * 32 * C::NQ3 + l0 + e; // intended group stride (benchmark05.cc:810-812)
            T x[R][NM];
            T *dst[R];
            bool ok[R];
#pragma unroll
            for (int k = 0; k < R; ++k)
            {
                ok[k]        = w + k * NW < NQ2;
                const int ji = ok[k] ? w + k * NW : 0;
                const int j = ji / NQ, i = ji - j * NQ;
                const T *sb[G]; // s2[r][j][i] sits at i ^ ((r*nm + j) & (G-1)) = i ^ ((r + j) & (G-1)): nm is odd
#pragma unroll
                for (int m = 0; m < G; ++m)
                    sb[m] = wk + (j * NQ + (i ^ ((j + m) & (G - 1)))) * EL + e;
#pragma unroll
                for (int r = 0; r < NM; ++r)
                    x[k][r] = sb[(r * NM) & (G - 1)][(r * NQ2) * EL];
                dst[k] = gout + 32 * ji;
            }
            coa_rows<T, NM, NQ, B2, R>(x, [&](int k, int kk, T v) { st_stream_if(dst[k] + (size_t)32 * NQ2 * kk, v, ok[k]); });
        }
    }
}

// Tried and dropped (tools/tune/lanes_probe.cu history, profiles/r02_coa_probe.csv): a two-slot input ring with one CTA
// per SM (0.61-0.69 at nq = 10 FP64); pass 1 with a barrier between its loads and its in-place stores (0.69); a q-outer
// form that fuses directions 0 and 1 in registers and reads the rows from the staged slot (39 % fewer shared-memory
// wavefronts, but 150 registers and 12 warps per SM: 0.66); 32-byte tiles (EL = 4 doubles: 0.60); the gather as one small
// bulk (TMA) copy per index of the tile (64 / 128 bytes each, issued by one warp, completion on an mbarrier) instead of
// cp.async: 3.4x SLOWER (0.21) -- the copy engine is made for kilobytes per instruction, not for 729 tiny ones.  And the
// other way round, the tiled-TMA gather in front of the lanes PLANE kernel (sumfac_lanes.cuh; boxes into the region that
// later holds t2, planes pulled into registers from there): 0.41-0.50 where the direct loads hold 0.73-0.95 -- that kernel
// wants its nm^2 loads per thread in flight from HBM, not a staged tile behind one more barrier.  What bounds this
// kernel at ~0.7 is the shared-memory data path next to the FP64 pipe: every 8-byte warp access is two wavefronts, a row
// costs 38 of them per 90 DFMAs (84 % of the LSU at FP64 peak), and the two pipes overlap only across warps.
} // namespace b200fe

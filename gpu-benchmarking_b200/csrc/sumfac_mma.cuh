// sumfac_mma.cuh -- "mma" back-end: FP64 sum-factorisation on the tensor cores
// (DMMA, mma.sync.m8n8k4.f64 = SASS DMMA.8x8x4) for the cases where the contraction is
// a dense GEMM-shaped problem that the scalar FP64 pipe cannot feed fast enough.  Quad:
// per element 2*nq*nm*(nm+nq) flops against 8*(nm^2+nq^2) bytes is 2.9-7.9 flop/B for
// nq = 12..32, i.e. 19-51 TFLOP/s at the HBM roofline, while DFMA with a constant-bank
// operand tops out near 27 TFLOP/s on B200 (profiles/r01_ubench_fp_pipe.txt) and DMMA
// sustains 37 TFLOP/s from registers with 1/8 of the issue slots
// (profiles/r01_ubench_dmma.txt).  Hex: nq = 8 fills the 8x8x4 tile exactly.
//
// Work unit: one WARP owns a group of G consecutive elements from load to store;
// warps never synchronise with each other (only __syncwarp), so their load /
// contract / store phases drift apart and overlap on the SM.
//
//   load   the group's contiguous slab in[e0 .. e0+G) is fetched by ONE bulk
//          tensor-memory-accelerator copy (cp.async.bulk, completion on the warp's
//          own mbarrier).  The copy of the next group is issued as soon as direction
//          0 has consumed the slot, so it lands while the later directions compute.
//   dir 0  mid[(e,..,q)][i] = sum_p in[(e,..,q)][p] * B0[p][i]
//          A = data rows (M = all rows of the group, flattened), B = basis fragments.
//   dir 1+ out[..][j][..]   = sum_q B1[q][j] * mid[..][q][..]
//          A = transposed basis fragments (M = j, K = q), B = the intermediate
//          (N = everything else, flattened); each lane ends up with two outputs
//          adjacent in i: one 16-byte store (streaming, to global, in the last direction).
//
// Padding of M/N/K up to the 8x8x4 tile is done with zeros in the basis
// fragments and with clamped / zero-selected data loads, never by touching
// memory outside the group.  DMMA.8x8x4 on sm_100 accumulates its four products
// in k order with fused multiply-adds, which is the reference's own summation
// order, so this back-end too is bit-identical to the reference kernels
// (tests/test_bwdtrans_gpu.py, tests/test_reference_kernels_gpu.py compare with
// array_equal); the documented bar, should hardware ever differ, is 1e-12.
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

// ---- the two pass shapes ------------------------------------------------------------------
//
// pass_data_rows: A = data.  nrows rows of NM values, row r at src + r*NM (contiguous rows: the raw
// element-major slab), contracted with the basis fragments fragB[ks][nt]; output (row, i) to
// dst + row*DS + i (shared memory).  MB m-tiles (8 rows each) share every B fragment.
// SWZ (NQ == 8 only): the destination rows are unpadded (DS == 8) and column i of row `row` is stored at
// i ^ (4 * ((row >> 1) & 1)): the double2 stores of a quarter warp (2 consecutive rows x 4 column pairs) and the scalar
// loads of the next pass's B fragments (4 consecutive rows x 4 columns per half warp) are both bank-conflict free, which
// no padding achieves (stride 12 served the loads and left the stores 2-way conflicted), and the tile is a third smaller.
template <int NQ, int DS, int MB, bool SWZ = false>
__device__ __forceinline__ void mma_pass_data_rows(const double *__restrict__ src, const double *__restrict__ fragB,
                                                   double *__restrict__ dst, int nrows, int lane)
{
    constexpr int NM = NQ - 1, KS = (NM + 3) / 4, NT = (NQ + 7) / 8;
    static_assert(!SWZ || (NQ == 8 && DS == 8), "the swizzle is defined for 8-wide unpadded rows");
    const int r = lane >> 2, c = lane & 3;
    const int cs = SWZ ? ((2 * c) ^ (4 * ((r >> 1) & 1))) : 2 * c; // 8 * tile + r keeps bit 1 of the row
#pragma unroll 1
    for (int mt = 0; mt * 8 < nrows; mt += MB)
    {
        double acc[MB][NT][2];
        const double *ap[MB];
#pragma unroll
        for (int m = 0; m < MB; ++m)
        {
            const int row = (mt + m) * 8 + r;
            ap[m]         = src + (row < nrows ? row : nrows - 1) * NM + c;
#pragma unroll
            for (int n = 0; n < NT; ++n)
                acc[m][n][0] = acc[m][n][1] = 0.0;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
        {
            double a[MB], b[NT];
#pragma unroll
            for (int m = 0; m < MB; ++m)
            {
                a[m] = ap[m][4 * ks]; // may over-read up to 3 values past the row: still inside the slot
                if (4 * ks + 3 >= NM && 4 * ks + c >= NM)
                    a[m] = 0.0; // k padding: whatever lies there (next row, Inf, NaN) must not contribute
            }
#pragma unroll
            for (int n = 0; n < NT; ++n)
                b[n] = fragB[(ks * NT + n) * 32 + lane];
#pragma unroll
            for (int m = 0; m < MB; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n)
                    dmma884(acc[m][n], a[m], b[n]);
        }
#pragma unroll
        for (int m = 0; m < MB; ++m)
        {
            const int row = (mt + m) * 8 + r;
            if (row < nrows)
            {
                double *mp = dst + row * DS + cs;
#pragma unroll
                for (int n = 0; n < NT; ++n)
                    if (8 * n + 2 * c < NQ)
                        *reinterpret_cast<double2 *>(mp + 8 * n) = make_double2(acc[m][n][0], acc[m][n][1]);
            }
        }
    }
}

// pass_basis_rows: A = transposed basis fragments fragA[mt][ks] (M = the NQ new points, K = NM).  The
// data is the B operand: column n of ncols decomposes as (g, w) = (n / W, n % W) and its K values lie
// STRIDE apart: src + (g*NM + k)*STRIDE + w.  Output (m, n) goes to dst + g*DG + m*DM + w, i.e. each
// lane holds two values adjacent in w: one 16-byte store, to shared memory or (streaming) to global.
template <int NQ, int W, int STRIDE, int DG, int DM, int NB, bool TO_GLOBAL, bool SUMSQ = false, bool SWZ_SRC = false>
__device__ __forceinline__ void mma_pass_basis_rows(const double *__restrict__ src, const double *__restrict__ fragA,
                                                    double *__restrict__ dst, int ncols, bool vec, int lane, double &ss)
{
    constexpr int NM = NQ - 1, KS = (NM + 3) / 4, MT = (NQ + 7) / 8;
    static_assert(W % 2 == 0, "outputs are paired along w");
    static_assert(!SWZ_SRC || (W == 8 && STRIDE == 8), "the swizzled source has 8-wide unpadded rows");
    const int r = lane >> 2, c = lane & 3;
#pragma unroll 1
    for (int nt = 0; nt * 8 < ncols; nt += NB)
    {
        double acc[MT][NB][2];
        const double *bp[NB];
#pragma unroll
        for (int t = 0; t < NB; ++t)
        {
            const int n  = (nt + t) * 8 + r;
            const int nc = n < ncols ? n : ncols - 1;
            const int g = nc / W, w = nc - g * W;
            bp[t] = src + (g * NM + c) * STRIDE + w;
            if (SWZ_SRC) // row g*NM + c + 4*ks: bit 1 of the row index picks the column half (4*ks keeps bit 1 only for even ks)
                bp[t] = src + (g * NM + c) * STRIDE; // the column is added per k step below
#pragma unroll
            for (int m = 0; m < MT; ++m)
                acc[m][t][0] = acc[m][t][1] = 0.0;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
        {
            double a[MT], b[NB];
#pragma unroll
            for (int m = 0; m < MT; ++m)
                a[m] = fragA[(m * KS + ks) * 32 + lane];
#pragma unroll
            for (int t = 0; t < NB; ++t)
            {
                int col = 0;
                if (SWZ_SRC)
                {
                    const int n  = (nt + t) * 8 + r;
                    const int nc = n < ncols ? n : ncols - 1;
                    const int g = nc / W, w = nc - g * W, row = g * NM + c + 4 * ks;
                    col = w ^ (4 * ((row >> 1) & 1));
                }
                if (4 * ks + 3 < NM)
                    b[t] = bp[t][4 * ks * STRIDE + col];
                else
                    b[t] = (4 * ks + c < NM) ? bp[t][4 * ks * STRIDE + col] : 0.0; // rows >= nm do not exist
            }
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int t = 0; t < NB; ++t)
                    dmma884(acc[m][t], a[m], b[t]);
        }
#pragma unroll
        for (int t = 0; t < NB; ++t)
        {
            const int n = (nt + t) * 8 + 2 * c;
            if (n < ncols)
            {
                const int g = n / W, w = n - g * W;
                double *op = dst + (size_t)g * DG + w;
#pragma unroll
                for (int m = 0; m < MT; ++m)
                {
                    const int j = 8 * m + r;
                    if (j < NQ)
                    {
                        if (SUMSQ) // fused checksum: this lane's share of sum(out^2), fixed order
                            ss = fmadd(acc[m][t][1], acc[m][t][1], fmadd(acc[m][t][0], acc[m][t][0], ss));
                        if (!TO_GLOBAL)
                            *reinterpret_cast<double2 *>(op + j * DM) = make_double2(acc[m][t][0], acc[m][t][1]);
                        else if (vec)
                            st_stream(reinterpret_cast<double2 *>(op + j * DM), make_double2(acc[m][t][0], acc[m][t][1]));
                        else
                        {
                            st_stream(op + j * DM, acc[m][t][0]);
                            st_stream(op + j * DM + 1, acc[m][t][1]);
                        }
                    }
                }
            }
        }
    }
}

// ---- fully unrolled twins for a FULL group (compile-time row / column counts) ------------------
// Every tile index is a compile-time constant, so all shared-memory offsets become immediates of
// one per-lane base register and no index arithmetic is left in the instruction stream: per element
// the generic loops issue ~15 instructions per DMMA (hex nq = 8), these ~3.
// The blocks are software pipelined by hand: the operand loads of block b+1 are issued before the
// DMMAs of block b, its stores after them.  dmma884_ordered keeps the issue order written here --
// all first k steps of a block's independent accumulators, then all second ones -- so that no DMMA
// waits for the one issued just before it (ptxas otherwise schedules each accumulator's dependent
// chain back to back).
__device__ __forceinline__ void dmma884_ordered(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// PRE: the basis fragments of this lane are handed over in registers (pre[ks*NT + n]) instead of being fetched from the
// fragment area at the start of every pass -- the kernels whose groups are always full (G = 1) load them once per warp
template <int NQ, int DS, int MB, int NROWS, bool SWZ = false, bool PRE = false>
__device__ __forceinline__ void mma_pass_data_rows_full(const double *__restrict__ src,
                                                        const double *__restrict__ fragB, double *__restrict__ dst,
                                                        int lane, const double *pre = nullptr)
{
    constexpr int NM = NQ - 1, KS = (NM + 3) / 4, NT = (NQ + 7) / 8, MT = (NROWS + 7) / 8;
    constexpr int NBLK = (MT + MB - 1) / MB;
    static_assert(!SWZ || (NQ == 8 && DS == 8), "the swizzle is defined for 8-wide unpadded rows");
    const int r = lane >> 2, c = lane & 3;
    const double *abase = src + r * NM + c;   // row (8*mt + r), column c
    double *dbase       = dst + r * DS + (SWZ ? ((2 * c) ^ (4 * ((r >> 1) & 1))) : 2 * c);
    const bool kpad     = c >= NM - 4 * (KS - 1); // this lane's column of the last k step is padding

    double b[KS][NT]; // the basis fragments are the same for every block: loaded once per pass
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int n = 0; n < NT; ++n)
            b[ks][n] = PRE ? pre[ks * NT + n] : fragB[(ks * NT + n) * 32 + lane];

    double a[2][MB][KS];
    auto load_block = [&](int blk, double (&dstA)[MB][KS]) {
#pragma unroll
        for (int m = 0; m < MB; ++m)
            if (blk * MB + m < MT)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                {
                    // rows >= NROWS of the last tile read whatever follows the slab inside this warp's
                    // region: rows are independent and those outputs are never stored
                    dstA[m][ks] = abase[(blk * MB + m) * 8 * NM + 4 * ks];
                    if (4 * ks + 3 >= NM && kpad)
                        dstA[m][ks] = 0.0; // k padding must not contribute whatever lies there
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
                const int mt     = blk * MB + m;
                const bool whole = mt * 8 + 8 <= NROWS;
                if (whole || r < NROWS - mt * 8)
                {
#pragma unroll
                    for (int n = 0; n < NT; ++n)
                        if (8 * n + 8 <= NQ || 8 * n + 2 * c < NQ)
                            *reinterpret_cast<double2 *>(dbase + mt * 8 * DS + 8 * n) =
                                make_double2(acc[m][n][0], acc[m][n][1]);
                }
            }
    }
}

// needs NCOLS % 8 == 0.  A tile of 8 columns may straddle two groups when W % 8 != 0: the lanes past
// the group boundary then add one compile-time constant to their address (a predicated add per tile).
template <int NQ, int W, int STRIDE, int DG, int DM, int NB, bool TO_GLOBAL, bool VEC, int NCOLS, bool SUMSQ = false,
          bool SWZ_SRC = false, bool PRE = false>
__device__ __forceinline__ void mma_pass_basis_rows_full(const double *__restrict__ src,
                                                         const double *__restrict__ fragA, double *__restrict__ dst,
                                                         int lane, double &ss, const double *pre = nullptr)
{
    constexpr int NM = NQ - 1, KS = (NM + 3) / 4, MT = (NQ + 7) / 8, NTT = NCOLS / 8;
    constexpr int NBLK = (NTT + NB - 1) / NB;
    static_assert(W % 2 == 0 && NCOLS % 8 == 0, "whole tiles, outputs paired along w");
    static_assert(!SWZ_SRC || (W == 8 && STRIDE == 8), "the swizzled source has 8-wide unpadded rows");
    const int r = lane >> 2, c = lane & 3;
    const double *bbase = src + c * STRIDE + r;    // k row c, column r of tile 0
    // swizzled source: this lane's k row is (compile-time base) + c and bit 1 of the row index picks the column half:
    // one per-lane base address for each residue of the compile-time base modulo 4
    const int t0 = (c >> 1) & 1, t1 = ((c + 1) >> 1) & 1;
    const double *bbs[4] = {src + c * STRIDE + (r ^ (4 * t0)), src + c * STRIDE + (r ^ (4 * t1)),
                            src + c * STRIDE + (r ^ (4 * (t0 ^ 1))), src + c * STRIDE + (r ^ (4 * (t1 ^ 1)))};
    double *dbase       = dst + r * DM + 2 * c;    // output row r, column pair c
    const bool kpad     = c >= NM - 4 * (KS - 1);

    double a[MT][KS]; // transposed basis fragments: loaded once per pass
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
            a[m][ks] = PRE ? pre[m * KS + ks] : fragA[(m * KS + ks) * 32 + lane];

    double b[2][NB][KS];
    auto load_block = [&](int blk, double (&dstB)[NB][KS]) {
#pragma unroll
        for (int t = 0; t < NB; ++t)
            if (blk * NB + t < NTT)
            {
                const int g = ((blk * NB + t) * 8) / W, w0 = ((blk * NB + t) * 8) % W;
                // columns w0 + r >= W belong to the next group: its rows start NM*STRIDE later
                const double *bb = (w0 + 8 > W && r >= W - w0) ? bbase + (NM * STRIDE - W) : bbase;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                {
                    if (SWZ_SRC) // W == 8: a tile never straddles groups
                        bb = bbs[(g * NM + 4 * ks) & 3];
                    if (4 * ks + 3 < NM)
                        dstB[t][ks] = bb[(g * NM + 4 * ks) * STRIDE + w0];
                    else
                        dstB[t][ks] = kpad ? 0.0 : bb[(g * NM + 4 * ks) * STRIDE + w0]; // rows >= nm do not exist
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
                const int g = ((blk * NB + t) * 8) / W, w0 = ((blk * NB + t) * 8) % W;
                double *db = (w0 + 8 > W && 2 * c >= W - w0) ? dbase + (DG - W) : dbase;
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    if (8 * m + 8 <= NQ || 8 * m + r < NQ)
                    {
                        double *op = db + (size_t)g * DG + w0 + 8 * m * DM;
                        if (SUMSQ)
                            ss = fmadd(acc[m][t][1], acc[m][t][1], fmadd(acc[m][t][0], acc[m][t][0], ss));
                        if (!TO_GLOBAL)
                            *reinterpret_cast<double2 *>(op) = make_double2(acc[m][t][0], acc[m][t][1]);
                        else if (VEC)
                            st_stream(reinterpret_cast<double2 *>(op), make_double2(acc[m][t][0], acc[m][t][1]));
                        else
                        {
                            st_stream(op, acc[m][t][0]);
                            st_stream(op + 1, acc[m][t][1]);
                        }
                    }
            }
    }
}

// dispatch: unrolled twin for a full group where the shape allows it, generic loops otherwise
template <int NQ, int DS, int MB, int NROWS, bool SWZ = false>
__device__ __forceinline__ void mma_dir_data(const double *__restrict__ src, const double *__restrict__ fragB,
                                             double *__restrict__ dst, int nrows, int lane)
{
    if (nrows == NROWS)
        mma_pass_data_rows_full<NQ, DS, MB, NROWS, SWZ>(src, fragB, dst, lane);
    else
        mma_pass_data_rows<NQ, DS, MB, SWZ>(src, fragB, dst, nrows, lane);
}
template <int NQ, int W, int STRIDE, int DG, int DM, int NB, bool TO_GLOBAL, int NCOLS, bool SUMSQ = false,
          bool SWZ_SRC = false>
__device__ __forceinline__ void mma_dir_basis(const double *__restrict__ src, const double *__restrict__ fragA,
                                              double *__restrict__ dst, int ncols, bool vec, int lane, double &ss)
{
    if constexpr (NCOLS % 8 == 0)
    {
        if (ncols == NCOLS)
        {
            if (vec)
                mma_pass_basis_rows_full<NQ, W, STRIDE, DG, DM, NB, TO_GLOBAL, true, NCOLS, SUMSQ, SWZ_SRC>(src, fragA, dst,
                                                                                                         lane, ss);
            else
                mma_pass_basis_rows_full<NQ, W, STRIDE, DG, DM, NB, TO_GLOBAL, false, NCOLS, SUMSQ, SWZ_SRC>(src, fragA, dst,
                                                                                                          lane, ss);
            return;
        }
    }
    mma_pass_basis_rows<NQ, W, STRIDE, DG, DM, NB, TO_GLOBAL, SUMSQ, SWZ_SRC>(src, fragA, dst, ncols, vec, lane, ss);
}

// this warp's share of the fused checksum -> partials[global warp index] (xor-shuffle: fixed order)
__device__ __forceinline__ void mma_store_partial(double ss, double *__restrict__ partials, int warp_global, int lane)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0)
        partials[warp_global] = ss;
}

// basis matrix B[p*NQ + i] -> B-operand fragments [ks][nt][lane] (K = p, N = i), zero padded
template <int NQ, int THREADS>
__device__ __forceinline__ void mma_fill_fragB(double *__restrict__ frag, const double *__restrict__ basis)
{
    constexpr int NM = NQ - 1, KS = (NM + 3) / 4, NT = (NQ + 7) / 8;
    for (int idx = threadIdx.x; idx < KS * NT * 32; idx += THREADS)
    {
        const int l = idx & 31, t = idx >> 5, nt = t % NT, ks = t / NT;
        const int p = 4 * ks + (l & 3), i = 8 * nt + (l >> 2);
        frag[idx]   = (p < NM && i < NQ) ? basis[p * NQ + i] : 0.0;
    }
}
// basis matrix B[q*NQ + j] -> A-operand fragments of its transpose [mt][ks][lane] (M = j, K = q)
template <int NQ, int THREADS>
__device__ __forceinline__ void mma_fill_fragA(double *__restrict__ frag, const double *__restrict__ basis)
{
    constexpr int NM = NQ - 1, KS = (NM + 3) / 4, MT = (NQ + 7) / 8;
    for (int idx = threadIdx.x; idx < MT * KS * 32; idx += THREADS)
    {
        const int l = idx & 31, t = idx >> 5, ks = t % KS, mt = t / KS;
        const int j = 8 * mt + (l >> 2), q = 4 * ks + (l & 3);
        frag[idx]   = (j < NQ && q < NM) ? basis[q * NQ + j] : 0.0;
    }
}

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

// ============================== quad ==========================================

template <int NQ, int G, int WARPS, int MB0, int NB1> struct QuadMma
{
    static constexpr int NM    = NQ - 1;
    static constexpr int NM2   = NM * NM;
    static constexpr int NQ2   = NQ * NQ;
    static constexpr int KS    = (NM + 3) / 4;     // k steps (K = nm in both directions)
    static constexpr int NT    = (NQ + 7) / 8;     // tiles over the nq new points of a direction
    static constexpr int S     = mma_mid_stride(NQ);
    static constexpr int SLOT  = (G * NM2 + 1 + 3 + 1) / 2 * 2; // +1: 8-byte window offset, +3: k over-read of the last row
    static constexpr int MID   = G * NM * S;
    static constexpr int WARP_D = SLOT + MID;      // doubles per warp (even)
    static constexpr int FRAG  = KS * NT * 32;
    static constexpr int BAR_BYTES = (WARPS * 8 + 15) / 16 * 16;
    static constexpr size_t SMEM = BAR_BYTES + (size_t)(2 * FRAG + WARPS * WARP_D) * sizeof(double);
    static_assert(NQ % 2 == 0, "the mma back-end pairs outputs along i");
};

// SUMSQ: also leave sum(out^2) of this warp's elements in partials[blockIdx.x * WARPS + warp] (SURVEY.md 8f-2:
// the checksum that follows every operator in the reference, fused into its epilogue instead of re-reading out)
template <int NQ, int G, int WARPS, int MB0, int NB1, bool SUMSQ = false>
__global__ void __launch_bounds__(WARPS * 32)
    bwdtrans_quad_mma_kernel(const double *__restrict__ basis0, const double *__restrict__ basis1,
                             const double *__restrict__ in, double *__restrict__ out, unsigned nelmt, unsigned ngroups,
                             int out_vec, double *__restrict__ partials)
{
    using C = QuadMma<NQ, G, WARPS, MB0, NB1>;
    constexpr int NM = C::NM, S = C::S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    double *frag0  = reinterpret_cast<double *>(smem_raw + C::BAR_BYTES);
    double *frag1  = frag0 + C::FRAG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *slot  = frag1 + C::FRAG + warp * C::WARP_D;
    double *mid   = slot + C::SLOT;
    uint64_t *bar = bars + warp;

    mma_fill_fragB<NQ, WARPS * 32>(frag0, basis0);
    mma_fill_fragA<NQ, WARPS * 32>(frag1, basis1);
    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // One input slot per warp: the next group's copy is issued as soon as direction 0 has drained it and
    // lands under direction 1.  (A two-slot ring was measured: the extra shared memory costs more resident
    // warps than the deeper prefetch gains -- profiles/r01_tune_mma.csv vs r01_tune_mma_2slot.csv.)
    const unsigned nw = gridDim.x * WARPS;
    unsigned g        = blockIdx.x * WARPS + warp;
    unsigned parity   = 0;
    double ss         = 0.0;   // fused checksum (SUMSQ)
    bool by_bar       = false; // the slot is filled through the mbarrier (else it was copied with plain loads)
    if (g < ngroups)
        by_bar = mma_fetch_group<G, C::NM2>(slot, bar, in, g, nelmt, lane);

    for (; g < ngroups; g += nw)
    {
        const size_t e0 = (size_t)g * G;
        const int ne    = (nelmt - e0 < (size_t)G) ? (int)(nelmt - e0) : G;
        const double *s_in =
            slot + ((reinterpret_cast<uintptr_t>(in + e0 * C::NM2) & 15u) >> 3); // window offset of this group
        if (by_bar)
        {
            mbar_wait(bar, parity);
            parity ^= 1u; // the phase advances only when the fill went through the barrier
        }
        // direction 0: mid[(e,q)][i] = sum_p in[(e,q)][p] B0[p][i]
        mma_dir_data<NQ, S, MB0, G * NM>(s_in, frag0, mid, ne * NM, lane);
        __syncwarp();
        // the slot is drained: start fetching this warp's next group under direction 1
        if (g + nw < ngroups)
            by_bar = mma_fetch_group<G, C::NM2>(slot, bar, in, g + nw, nelmt, lane);
        // direction 1: out[e][j][i] = sum_q B1[q][j] mid[e][q][i]
        mma_dir_basis<NQ, NQ, S, C::NQ2, NQ, NB1, true, G * NQ, SUMSQ>(mid, frag1, out + e0 * C::NQ2, ne * NQ,
                                                                      out_vec != 0, lane, ss);
        __syncwarp(); // mid is rewritten by the next group's direction 0
    }
    if (SUMSQ)
        mma_store_partial(ss, partials, blockIdx.x * WARPS + warp, lane);
}

// ============================== hex ===========================================

// stride of the (e, r) rows of the second intermediate: >= nq^2, even, == 4 or 12 (mod 16)
constexpr int mma_s2_stride(int nq)
{
    int s = nq * nq;
    while (s % 16 != 4 && s % 16 != 12)
        s += 2;
    return s;
}

template <int NQ, int G, int WARPS, int MB0, int NB> struct HexMma
{
    static constexpr int NM   = NQ - 1;
    static constexpr int NM2  = NM * NM;
    static constexpr int NM3  = NM2 * NM;
    static constexpr int NQ2  = NQ * NQ;
    static constexpr int NQ3  = NQ2 * NQ;
    static constexpr int KS   = (NM + 3) / 4;
    static constexpr int NT   = (NQ + 7) / 8;
    static constexpr bool SWZ = NQ == 8;             // 8-wide rows: unpadded + XOR swizzle instead of padding (mma_pass_data_rows)
    static constexpr int S1   = SWZ ? NQ : mma_mid_stride(NQ); // s1[(e,r,q)][i]
    static constexpr int S2   = mma_s2_stride(NQ);  // s2[(e,r)][j*nq + i]
    static constexpr int SLOT = (G * NM3 + 1 + 3 + 1) / 2 * 2;
    static constexpr int MID1 = G * NM2 * S1;
    static constexpr int MID2 = G * NM * S2;
    static constexpr int WARP_D = SLOT + MID1 + MID2;
    static constexpr int FRAG = KS * NT * 32;
    static constexpr int BAR_BYTES = (WARPS * 8 + 15) / 16 * 16;
    static constexpr size_t SMEM = BAR_BYTES + (size_t)(3 * FRAG + WARPS * WARP_D) * sizeof(double);
    static_assert(NQ % 2 == 0, "the mma back-end pairs outputs along i");
};

template <int NQ, int G, int WARPS, int MB0, int NB, bool SUMSQ = false>
__global__ void __launch_bounds__(WARPS * 32)
    bwdtrans_hex_mma_kernel(const double *__restrict__ basis0, const double *__restrict__ basis1,
                            const double *__restrict__ basis2, const double *__restrict__ in, double *__restrict__ out,
                            unsigned nelmt, unsigned ngroups, int out_vec, double *__restrict__ partials)
{
    using C = HexMma<NQ, G, WARPS, MB0, NB>;
    constexpr int NM = C::NM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    double *frag0  = reinterpret_cast<double *>(smem_raw + C::BAR_BYTES);
    double *frag1  = frag0 + C::FRAG;
    double *frag2  = frag1 + C::FRAG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *slot  = frag2 + C::FRAG + warp * C::WARP_D;
    double *s1    = slot + C::SLOT;
    double *s2    = s1 + C::MID1;
    uint64_t *bar = bars + warp;

    mma_fill_fragB<NQ, WARPS * 32>(frag0, basis0);
    mma_fill_fragA<NQ, WARPS * 32>(frag1, basis1);
    mma_fill_fragA<NQ, WARPS * 32>(frag2, basis2);
    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const unsigned nw = gridDim.x * WARPS;
    unsigned g        = blockIdx.x * WARPS + warp;
    unsigned parity   = 0;
    double ss         = 0.0;
    bool by_bar       = false;
    if (g < ngroups)
        by_bar = mma_fetch_group<G, C::NM3>(slot, bar, in, g, nelmt, lane);

    // G = 1: every group is a whole element, so every pass is the unrolled one and this lane's basis fragments never
    // change: they live in registers for the whole kernel instead of being fetched from the fragment area by every pass
    // of every element (6 of ~100 shared-memory instructions per element; the L1 data pipe is what bounds this kernel)
    constexpr bool PRE = G == 1 && (G * C::NM * NQ) % 8 == 0 && (G * C::NQ2) % 8 == 0;
    constexpr int MT   = (NQ + 7) / 8;
    double p0[PRE ? C::KS * C::NT : 1], p1[PRE ? MT * C::KS : 1], p2[PRE ? MT * C::KS : 1];
    if constexpr (PRE)
    {
#pragma unroll
        for (int f = 0; f < C::KS * C::NT; ++f)
            p0[f] = frag0[f * 32 + lane];
#pragma unroll
        for (int f = 0; f < MT * C::KS; ++f)
        {
            p1[f] = frag1[f * 32 + lane];
            p2[f] = frag2[f * 32 + lane];
        }
    }

    for (; g < ngroups; g += nw)
    {
        const size_t e0 = (size_t)g * G;
        const int ne    = (nelmt - e0 < (size_t)G) ? (int)(nelmt - e0) : G;
        const double *s_in = slot + ((reinterpret_cast<uintptr_t>(in + e0 * C::NM3) & 15u) >> 3);
        if (by_bar)
        {
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        // direction 0: s1[(e,r,q)][i] = sum_p in[(e,r,q)][p] B0[p][i]
        if constexpr (PRE)
            mma_pass_data_rows_full<NQ, C::S1, MB0, G * C::NM2, C::SWZ, true>(s_in, frag0, s1, lane, p0);
        else
            mma_dir_data<NQ, C::S1, MB0, G * C::NM2, C::SWZ>(s_in, frag0, s1, ne * C::NM2, lane);
        __syncwarp();
        if (g + nw < ngroups)
            by_bar = mma_fetch_group<G, C::NM3>(slot, bar, in, g + nw, nelmt, lane);
        // direction 1: s2[(e,r)][j][i] = sum_q B1[q][j] s1[(e,r)][q][i]
        if constexpr (PRE)
            mma_pass_basis_rows_full<NQ, NQ, C::S1, C::S2, NQ, NB, false, true, G * NM * NQ, false, C::SWZ, true>(s1, frag1, s2,
                                                                                                             lane, ss, p1);
        else
            mma_dir_basis<NQ, NQ, C::S1, C::S2, NQ, NB, false, G * NM * NQ, false, C::SWZ>(s1, frag1, s2, ne * NM * NQ, true,
                                                                                          lane, ss);
        __syncwarp();
        // direction 2: out[e][k][(j,i)] = sum_r B2[r][k] s2[e][r][(j,i)]
        if constexpr (PRE)
        {
            if (out_vec != 0)
                mma_pass_basis_rows_full<NQ, C::NQ2, C::S2, C::NQ3, C::NQ2, NB, true, true, G * C::NQ2, SUMSQ, false, true>(
                    s2, frag2, out + e0 * C::NQ3, lane, ss, p2);
            else
                mma_pass_basis_rows_full<NQ, C::NQ2, C::S2, C::NQ3, C::NQ2, NB, true, false, G * C::NQ2, SUMSQ, false, true>(
                    s2, frag2, out + e0 * C::NQ3, lane, ss, p2);
        }
        else
            mma_dir_basis<NQ, C::NQ2, C::S2, C::NQ3, C::NQ2, NB, true, G * C::NQ2, SUMSQ>(
                s2, frag2, out + e0 * C::NQ3, ne * C::NQ2, out_vec != 0, lane, ss);
        __syncwarp(); // s1 / s2 are rewritten by the next group
    }
    if (SUMSQ)
        mma_store_partial(ss, partials, blockIdx.x * WARPS + warp, lane);
}

} // namespace b200fe

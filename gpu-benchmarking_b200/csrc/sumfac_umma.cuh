// sumfac_umma.cuh -- FP32 quad BwdTrans at nq = 32 on the 5th-generation tensor cores (tcgen05.mma kind::tf32,
// operands and accumulators in TMEM), the one FP32 case of the BASELINE sweeps that is compute-bound on every other
// path (103 TFLOP/s FP32 needed at the HBM roofline; mma.sync 3xTF32 reaches 0.55 of it, FFMA2 0.3).
//
//   out[e][j][i] = sum_q ( sum_p in[e][q][p] B0[p][i] ) B1[q][j]            (benchmark04.cc:49-72), nm = 31, nq = 32
//
// as two tensor-core contractions per tile of 4 elements (M = 128 rows), with the 3xTF32 split (x = hi + lo, both
// TF32; hi*hi + hi*lo + lo*hi, FP32 accumulation) folded into the operand shapes:
//
//   direction 0   D0[(e,q)][n] = sum_p A0[(e,q)][p] * B0P[n][p]     B0P = [B0hi | B0lo] (N = 64, shared memory)
//                 A0hi x B0P (N = 64) then A0lo x B0hi (N = 32, accumulated onto columns 0..31);
//                 A0 is an operand IN TENSOR MEMORY: the thread that owns row (e,q) writes its 32 hi and 32 lo values
//                 into its own TMEM lane (tcgen05.st), so direction 0 reads only the 8 KB basis operand from shared
//                 memory;  T1[(e,q)][i] = D0[.][i] + D0[.][32 + i]
//   direction 1   D1[(e,i)][n] = sum_q A1[(e,i)][q] * B1P[n][q]     A1[(e,i)][q] = T1[(e,q)][i]: exactly what a thread
//                 reads back from TMEM lane (e,q) -- 32 consecutive i -- so it is stored to shared memory as an
//                 MN-MAJOR operand (no transpose);  out[e][j][i] = D1[(e,i)][j] + D1[(e,i)][32 + j], lanes = i:
//                 coalesced stores.  (A TMEM operand must be K-major, which here would need the transpose.)
//
// One persistent CTA per SM, 14 warps in five roles connected by mbarrier rings:
//   producer (1 thread)   bulk (TMA) copy of the tile's contiguous 15 376-byte slab, 4 slabs deep   -> RAW[rs]
//   convert  (4 warps)    thread = row (e,q): 31 values -> hi / lo -> tcgen05.st                    -> TMEM A0H[s], A0L[s]
//   mma      (1 thread)   8 + 8 tcgen05.mma per tile, tcgen05.commit onto the rings                 -> TMEM D0[s], D1[s]
//   epilogue0 (4 warps)   tcgen05.ld lane (e,q): T1 row -> hi / lo -> SW128_32B MN-major            -> smem A1H[s], A1L[s]
//   epilogue1 (4 warps)   tcgen05.ld lane (e,i): 32 outputs -> 32 coalesced streaming stores (+ fused checksum)
// The MMA thread issues direction 0 of tile t+1 before direction 1 of tile t, so the tensor core never waits for
// epilogue0's round trip.  Shared-memory traffic is what bounds this kernel (the first version, with A0 in shared memory
// too, spent 1300 of its 2100 clk per tile issuing MMAs that waited for operand reads); per tile it is now 31 KB for
// the slab, 32 KB of A1 stores (4 wavefronts per warp-wide 16-byte store: the chunk order is rotated with bit 2 of
// the lane) and ~50 KB of operand reads.  Descriptor encodings and the MN-major layout were validated in isolation
// first (tools/ubench/umma_tf32.cu -> profiles/r02_ubench_umma_tf32.txt).  Agrees with the reference's FFMA chain to
// rounding (component-wise 1e-5 bound, measured 2.5e-7; include/b200fe.h).
#pragma once

#include "common.cuh"
#include "sumfac_rows.cuh" // mbarrier / bulk-copy helpers

namespace b200fe
{
namespace umma
{

constexpr int NQ = 32, NM = 31, NM2 = NM * NM, NQ2 = NQ * NQ;
constexpr int TILE_E      = 4;                    // elements per tile: 4 x 32 rows = M = 128
constexpr int STAGES      = 2;                    // TMEM operand / accumulator sets and A1 tiles
constexpr int RAW_STAGES  = 4;                    // input slabs in flight
constexpr int RAW_BYTES   = TILE_E * NM2 * 4;     // 15 376, a multiple of 16
constexpr int RAW_STRIDE  = 15488;                // padded to a multiple of 128
constexpr int OPND_BYTES  = 128 * 128;            // one 128 x 32 tf32 operand tile
constexpr int BOP_BYTES   = 64 * 128;             // [hi | lo] basis operand, 64 rows of 32 k
constexpr int OFF_A1H = 0, OFF_A1L = OPND_BYTES;
constexpr int STAGE_BYTES = 2 * OPND_BYTES;       // 32 768
constexpr int OFF_STAGES  = 0;                                   // 1024-byte aligned operand tiles first
constexpr int OFF_B0P     = OFF_STAGES + STAGES * STAGE_BYTES;   // 65 536
constexpr int OFF_B1P     = OFF_B0P + BOP_BYTES;
constexpr int OFF_RAW     = OFF_B1P + BOP_BYTES;                 // 81 920
constexpr int OFF_BARS    = OFF_RAW + RAW_STAGES * RAW_STRIDE;   // 143 872
constexpr int NBARS       = 2 * RAW_STAGES + 8 * STAGES;
constexpr int OFF_TMEM    = OFF_BARS + NBARS * 8;
constexpr size_t SMEM     = OFF_TMEM + 16 + 1024;                // + slack to align the base to 1024 bytes
constexpr int THREADS     = 14 * 32;
// tensor memory, per stage: A0hi 32 | A0lo 32 | D0 64 | D1 64 columns
constexpr int TM_STAGE = 192, TM_A0H = 0, TM_A0L = 32, TM_D0 = 64, TM_D1 = 128;
constexpr int TMEM_COLS = 512; // power of two >= STAGES * TM_STAGE

// barrier indices: the RAW ring first, then the per-stage ones
constexpr int RAW_FULL = 0, RAW_EMPTY = RAW_STAGES, PER_STAGE = 2 * RAW_STAGES;
enum Bar
{
    A0_FULL = 0, A0_EMPTY, D0_FULL, D0_EMPTY, A1_FULL, A1_EMPTY, D1_FULL, D1_EMPTY
};

__device__ __forceinline__ uint32_t s32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
// exactly one lane of the (converged) warp gets true.  Unlike `lane == 0` the compiler knows that a single thread is
// active behind it, so the uniform-register operands of tcgen05.mma / tcgen05.commit are moved with one R2UR each;
// behind `lane == 0` every instruction was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~80 clk per MMA issued,
// more than the tensor core needs to execute it)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t saddr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts_f32x4(uint32_t saddr, const float4 &v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tc_fence_before()
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after()
{
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, one K = 8 step
__device__ __forceinline__ void tc_mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// D[tmem] (+)= A[tmem: lanes = rows, 8 columns = K] * B[smem desc]^T
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                 "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_ld_wait()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
                 "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
                 : "memory");
}
__device__ __forceinline__ void tc_st_wait()
{
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor: start address >> 4 in [0,14), leading byte offset >> 4
// in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout type in [61,64))
// K-major, 128-byte swizzle: rows of 32 tf32, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major tf32: SWIZZLE_128B_BASE32B (the only layout the tensor core takes for 4-byte MN-major operands): 32
// consecutive m (128 B) per k row, 4-row atoms (512 B), 32-byte chunk j of k row r at position j ^ (r % 4);
// 32-wide m blocks 4096 bytes apart (LBO), 4-deep k groups 512 bytes apart (SBO)
__device__ __forceinline__ uint64_t desc_mn_sw128_32b(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(4096u >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6), A = B = TF32 [7,10) [10,13), A major [15],
// B major [16] (0 = K), N >> 3 in [17,23), M >> 4 in [24,29)
constexpr uint32_t idesc_tf32(int m, int n, int a_mn_major)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

// x = hi + lo with hi the nearest TF32 (10 mantissa bits; integer rounding of the bit pattern) and lo the exact FP32
// remainder, of which the tensor core uses the leading 10 bits
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
}

// the [hi | lo] basis operand: row n < 32 holds hi(B[k][n]), row 32 + n holds lo(B[k][n]), k = 0..30, k = 31 zero;
// K-major SW128
__device__ __forceinline__ void fill_basis_operand(unsigned char *dst, const float *__restrict__ basis, int tid)
{
    for (int t = tid; t < 32 * 32; t += THREADS)
    {
        const int n = t >> 5, k = t & 31; // consecutive threads: consecutive k of one output column n
        float hi = 0.f, lo = 0.f;
        if (k < NM)
            split_tf32(basis[k * NQ + n], hi, lo);
        const int off = (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 2) ^ (n & 7)) << 4) | ((k & 3) << 2));
        *reinterpret_cast<float *>(dst + off)            = hi;
        *reinterpret_cast<float *>(dst + 4 * 1024 + off) = lo; // rows 32..63: four 8-row atoms further
    }
}

struct Phase // parity bookkeeping of one role on a ring: the k-th use of stage s waits on parity (k & 1)
{
    unsigned bits = 0;
    __device__ __forceinline__ unsigned full(int s) const { return (bits >> s) & 1u; }          // consumer side
    __device__ __forceinline__ unsigned empty(int s) const { return ((bits >> s) & 1u) ^ 1u; }  // producer side
    __device__ __forceinline__ void advance(int s) { bits ^= 1u << s; }
};

// PROF (development only, tools/umma_check.py): lane 0 of one warp per role of CTA 0 accumulates the cycles it spends
// in each barrier wait and in the whole loop into prof[role * 8 + k]
template <bool PROF> struct RoleTimer
{
    long long t0 = 0, acc[6] = {0, 0, 0, 0, 0, 0};
    __device__ __forceinline__ void start()
    {
        if (PROF)
            t0 = clock64();
    }
    __device__ __forceinline__ void lap(int k) // cycles since the last start() / lap() go to slot k
    {
        if (PROF)
        {
            const long long t = clock64();
            acc[k] += t - t0;
            t0 = t;
        }
    }
    __device__ __forceinline__ void dump(unsigned long long *prof, int role, int lane) const
    {
        if (PROF && prof && blockIdx.x == 0 && lane == 0)
            for (int k = 0; k < 6; ++k)
                prof[role * 8 + k] = (unsigned long long)acc[k];
    }
};

template <bool SUMSQ, bool PROF = false>
__global__ void __launch_bounds__(THREADS, 1)
    bwdtrans_quad32_umma_kernel(const float *__restrict__ basis0, const float *__restrict__ basis1,
                                const float *__restrict__ in, float *__restrict__ out, unsigned nelmt, unsigned ntiles,
                                double *__restrict__ partials, unsigned long long *__restrict__ prof)
{
    RoleTimer<PROF> tm;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023u) & ~uintptr_t(1023));
    uint64_t *bars      = reinterpret_cast<uint64_t *>(smem + OFF_BARS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_TMEM);
    const uint32_t sbase = s32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar  = [&](int which, int s) { return bars + PER_STAGE + which * STAGES + s; };
    auto rbar = [&](int which, int rs) { return bars + which + rs; }; // which = RAW_FULL / RAW_EMPTY

    fill_basis_operand(smem + OFF_B0P, basis0, tid);
    fill_basis_operand(smem + OFF_B1P, basis1, tid);
    if (tid == 0)
    {
        for (int rs = 0; rs < RAW_STAGES; ++rs)
        {
            mbar_init(rbar(RAW_FULL, rs), 1);
            mbar_init(rbar(RAW_EMPTY, rs), 4);
        }
        for (int s = 0; s < STAGES; ++s)
        {
            mbar_init(bar(A0_FULL, s), 4);
            mbar_init(bar(A0_EMPTY, s), 1);
            mbar_init(bar(D0_FULL, s), 1);
            mbar_init(bar(D0_EMPTY, s), 4);
            mbar_init(bar(A1_FULL, s), 4);
            mbar_init(bar(A1_EMPTY, s), 1);
            mbar_init(bar(D1_FULL, s), 1);
            mbar_init(bar(D1_EMPTY, s), 4);
        }
        mbar_fence_init();
    }
    if (warp == 13)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async(); // the basis operands were written through the generic proxy; the tensor core reads them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4)
    {
        // ---- convert: RAW slab -> A0hi / A0lo in this thread's TMEM lane; thread = row (e, q) of the tile -----------
        const int e = warp, q = lane;
        const uint32_t tlane = tmem + ((uint32_t)(e * 32) << 16);
        Phase ph, phr;
        unsigned it = 0;
        for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it)
        {
            const int s = it & 1, rs = it % RAW_STAGES;
            const size_t e0   = (size_t)tile * TILE_E;
            const int ne      = (nelmt - e0 < (size_t)TILE_E) ? (int)(nelmt - e0) : TILE_E;
            const bool staged = ne == TILE_E; // a ragged last tile is read straight from global memory
            float x[32];
            tm.start();
            mbar_wait(rbar(RAW_FULL, rs), phr.full(rs));
            tm.lap(0);
#pragma unroll
            for (int p = 0; p < 32; ++p)
                x[p] = 0.f;
            if (e < ne && q < NM)
            {
                if (staged)
                {
                    const uint32_t src = sbase + OFF_RAW + rs * RAW_STRIDE + (e * NM2 + q * NM) * 4;
#pragma unroll
                    for (int p = 0; p < NM; ++p)
                        x[p] = lds_f32(src + 4 * p); // lanes = rows 31 words apart: conflict-free
                }
                else
                {
                    const float *src = in + (e0 + e) * NM2 + q * NM;
#pragma unroll
                    for (int p = 0; p < NM; ++p)
                        x[p] = src[p];
                }
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(rbar(RAW_EMPTY, rs)); // the slab is in registers
            phr.advance(rs);
            tm.lap(1);
            mbar_wait(bar(A0_EMPTY, s), ph.empty(s)); // direction 0 of the tile two back has read A0[s]
            tc_fence_after();
            tm.lap(2);
            const uint32_t ta = tlane + (uint32_t)(s * TM_STAGE);
#pragma unroll
            for (int half = 0; half < 2; ++half)
            {
                float h[16], l[16];
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    split_tf32(x[16 * half + k], h[k], l[k]);
                tc_st16(ta + TM_A0H + 16 * half, h);
                tc_st16(ta + TM_A0L + 16 * half, l);
            }
            tc_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(bar(A0_FULL, s));
            ph.advance(s);
            tm.lap(3);
        }
        if (warp == 0)
            tm.dump(prof, 0, lane);
    }
    else if (warp < 8)
    {
        // ---- epilogue 0: D0 (TMEM lane (e,q), 64 columns) -> T1 row -> A1hi / A1lo, MN-major ----------------------
        const int e = warp & 3, q = lane;
        const bool swap = (q >> 2) & 1; // rotates the order of each chunk pair: 8 distinct bank groups per store
        Phase ph;
        unsigned it = 0;
        for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it)
        {
            const int s          = it & 1;
            const uint32_t taddr = tmem + (uint32_t)(s * TM_STAGE + TM_D0) + ((uint32_t)(e * 32) << 16);
            tm.start();
            mbar_wait(bar(D0_FULL, s), ph.full(s));
            tc_fence_after();
            tm.lap(0);
            mbar_wait(bar(A1_EMPTY, s), ph.empty(s)); // direction 1 of the tile two back has read A1[s]
            tm.lap(1);
            const uint32_t row = sbase + OFF_STAGES + s * STAGE_BYTES + e * 4096 + q * 128;
#pragma unroll
            for (int half = 0; half < 2; ++half)
            {
                uint32_t a[16], b[16];
                tc_ld16(taddr + half * 16, a);
                tc_ld16(taddr + 32 + half * 16, b);
                tc_ld_wait();
                if (half == 1)
                {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(bar(D0_EMPTY, s)); // D0[s] is in registers: direction 0 of tile t + 2 may overwrite it
                }
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) // pairs of 16-byte chunks: i = 16*half + 8*pr + (0..3 | 4..7)
                {
                    float4 h0, l0, h1, l1;
                    const int o = 8 * pr;
                    split_tf32(__uint_as_float(a[o + 0]) + __uint_as_float(b[o + 0]), h0.x, l0.x);
                    split_tf32(__uint_as_float(a[o + 1]) + __uint_as_float(b[o + 1]), h0.y, l0.y);
                    split_tf32(__uint_as_float(a[o + 2]) + __uint_as_float(b[o + 2]), h0.z, l0.z);
                    split_tf32(__uint_as_float(a[o + 3]) + __uint_as_float(b[o + 3]), h0.w, l0.w);
                    split_tf32(__uint_as_float(a[o + 4]) + __uint_as_float(b[o + 4]), h1.x, l1.x);
                    split_tf32(__uint_as_float(a[o + 5]) + __uint_as_float(b[o + 5]), h1.y, l1.y);
                    split_tf32(__uint_as_float(a[o + 6]) + __uint_as_float(b[o + 6]), h1.z, l1.z);
                    split_tf32(__uint_as_float(a[o + 7]) + __uint_as_float(b[o + 7]), h1.w, l1.w);
                    // the pair is one 32-byte chunk j of the 128-byte k row, stored at chunk position j ^ (q % 4); lanes
                    // with bit 2 of q set store its upper half first: a warp-wide store then touches 8 distinct
                    // 16-byte bank groups (4 wavefronts) instead of 4 (8 wavefronts)
                    const int j        = half * 2 + pr;
                    const uint32_t c32 = row + (uint32_t)((j ^ (q & 3)) << 5);
                    const uint32_t first = c32 + (swap ? 16u : 0u), second = c32 + (swap ? 0u : 16u);
                    const float4 hf = swap ? h1 : h0, hs = swap ? h0 : h1, lf = swap ? l1 : l0, ls = swap ? l0 : l1;
                    sts_f32x4(first + OFF_A1H, hf);
                    sts_f32x4(first + OFF_A1L, lf);
                    sts_f32x4(second + OFF_A1H, hs);
                    sts_f32x4(second + OFF_A1L, ls);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(bar(A1_FULL, s));
            ph.advance(s);
            tm.lap(2);
        }
        if (warp == 4)
            tm.dump(prof, 1, lane);
    }
    else if (warp < 12)
    {
        // ---- epilogue 1: D1 (TMEM lane (e,i), 64 columns) -> out[e][j][i], lanes = i ------------------------------
        const int e = warp & 3, i = lane;
        Phase ph;
        double ss   = 0.0;
        unsigned it = 0;
        for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it)
        {
            const int s          = it & 1;
            const size_t e0      = (size_t)tile * TILE_E;
            const int ne         = (nelmt - e0 < (size_t)TILE_E) ? (int)(nelmt - e0) : TILE_E;
            const uint32_t taddr = tmem + (uint32_t)(s * TM_STAGE + TM_D1) + ((uint32_t)(e * 32) << 16);
            float *dst           = out + (e0 + e) * NQ2 + i;
            tm.start();
            mbar_wait(bar(D1_FULL, s), ph.full(s));
            tc_fence_after();
            tm.lap(0);
#pragma unroll
            for (int half = 0; half < 2; ++half)
            {
                uint32_t a[16], b[16];
                tc_ld16(taddr + half * 16, a);
                tc_ld16(taddr + 32 + half * 16, b);
                tc_ld_wait();
                if (half == 1)
                {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(bar(D1_EMPTY, s));
                    tm.lap(1);
                }
                if (e < ne)
                {
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                    {
                        const float v = __uint_as_float(a[c]) + __uint_as_float(b[c]);
                        st_stream(dst + (half * 16 + c) * NQ, v);
                        if (SUMSQ)
                            ss = fma((double)v, (double)v, ss);
                    }
                }
            }
            ph.advance(s);
            tm.lap(2);
        }
        if (warp == 8)
            tm.dump(prof, 2, lane);
        if (SUMSQ)
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0)
                partials[blockIdx.x * 4 + e] = ss;
        }
    }
    else if (warp == 12)
    {
        // ---- producer: one bulk copy per tile, RAW_STAGES slabs ahead -----------------------------------------------
        if (lane == 0)
        {
            Phase phr;
            unsigned it = 0;
            long long c_begin = 0;
            unsigned long long ns_begin = 0;
            if (PROF)
            {
                c_begin = clock64();
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
            }
            for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it)
            {
                const int rs    = it % RAW_STAGES;
                const size_t e0 = (size_t)tile * TILE_E;
                const bool full = nelmt - e0 >= (size_t)TILE_E;
                tm.start();
                mbar_wait(rbar(RAW_EMPTY, rs), phr.empty(rs));
                tm.lap(0);
                if (full)
                {
                    fence_proxy_async(); // the converters' reads of RAW[rs] are ordered before the async write
                    mbar_arrive_expect_tx(rbar(RAW_FULL, rs), RAW_BYTES);
                    bulk_load(smem + OFF_RAW + rs * RAW_STRIDE, in + e0 * NM2, RAW_BYTES, rbar(RAW_FULL, rs));
                }
                else
                    mbar_arrive(rbar(RAW_FULL, rs)); // ragged tile: the converters read global memory themselves
                phr.advance(rs);
                tm.lap(1);
            }
            if (PROF) // SM clocks and wall nanoseconds over the producer's loop: the clock the SM really ran at
            {
                unsigned long long ns_end;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_end));
                tm.acc[4] = clock64() - c_begin;
                tm.acc[5] = (long long)(ns_end - ns_begin);
            }
            tm.dump(prof, 3, lane);
        }
        __syncwarp();
    }
    else
    {
        // ---- mma: the whole warp runs the loop (converged), one elected lane issues; direction 0 of tile t + 1 is issued
        // before direction 1 of tile t
        constexpr uint32_t I0_64 = idesc_tf32(128, 64, 0), I0_32 = idesc_tf32(128, 32, 0);
        constexpr uint32_t I1_64 = idesc_tf32(128, 64, 1), I1_32 = idesc_tf32(128, 32, 1);
        const uint64_t b0 = desc_k_sw128(sbase + OFF_B0P), b1 = desc_k_sw128(sbase + OFF_B1P);
        Phase ph0, ph1;
        const unsigned my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
        auto dir0 = [&](unsigned it) {
            const int s       = it & 1;
            const uint32_t ts = tmem + (uint32_t)(s * TM_STAGE);
            tm.start();
            mbar_wait(bar(A0_FULL, s), ph0.full(s));
            tm.lap(0);
            mbar_wait(bar(D0_EMPTY, s), ph0.empty(s));
            tm.lap(1);
            tc_fence_after();
            if (elect_one())
            {
#pragma unroll
                for (int k = 0; k < 4; ++k) // K = 8: 8 TMEM columns of A, 32 bytes of the K-major B rows per step
                    tc_mma_tf32_ts(ts + TM_D0, ts + TM_A0H + 8 * k, b0 + 2 * k, I0_64, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_tf32_ts(ts + TM_D0, ts + TM_A0L + 8 * k, b0 + 2 * k, I0_32, 1u);
                tc_commit(bar(D0_FULL, s));
                tc_commit(bar(A0_EMPTY, s));
            }
            __syncwarp();
            ph0.advance(s);
            tm.lap(4);
        };
        auto dir1 = [&](unsigned it) {
            const int s         = it & 1;
            const uint32_t base = sbase + OFF_STAGES + s * STAGE_BYTES;
            const uint64_t ah = desc_mn_sw128_32b(base + OFF_A1H), al = desc_mn_sw128_32b(base + OFF_A1L);
            const uint32_t d1 = tmem + (uint32_t)(s * TM_STAGE + TM_D1);
            tm.start();
            mbar_wait(bar(A1_FULL, s), ph1.full(s));
            tm.lap(2);
            mbar_wait(bar(D1_EMPTY, s), ph1.empty(s));
            tm.lap(3);
            tc_fence_after();
            if (elect_one())
            {
#pragma unroll
                for (int k = 0; k < 4; ++k) // K = 8: two 4-deep k groups = 1024 bytes per step
                    tc_mma_tf32_ss(d1, ah + 64 * k, b1 + 2 * k, I1_64, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_tf32_ss(d1, al + 64 * k, b1 + 2 * k, I1_32, 1u);
                tc_commit(bar(D1_FULL, s));
                tc_commit(bar(A1_EMPTY, s));
            }
            __syncwarp();
            ph1.advance(s);
            tm.lap(5);
        };
        if (my_tiles)
            dir0(0);
        for (unsigned it = 0; it < my_tiles; ++it)
        {
            if (it + 1 < my_tiles)
                dir0(it + 1);
            dir1(it);
        }
        tm.dump(prof, 4, lane);
    }

    tc_fence_before();
    __syncthreads(); // every role has left its loop: all tcgen05 operations of this CTA were waited for by their readers
    if (warp == 13)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

} // namespace umma
} // namespace b200fe

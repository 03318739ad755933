// umma_tile.cu -- tcgen05.mma kind::tf32 instruction streams in isolation (one issuing thread per SM, no other warps,
// no data movement): what does the tensor core itself need for the instruction mix of csrc/sumfac_umma.cuh?
// A stream = G groups of 4 K-steps (M = 128, K = 8 each), each group with its own A kind (K-major smem, MN-major smem,
// TMEM), N, accumulator column and accumulate-from-first flag; `ncommit` commits after every `per` groups.
//   nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -o umma_tile umma_tile.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k(uint32_t a) { return (uint64_t)((a & 0x3ffffu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t a) { return (uint64_t)((a & 0x3ffffu) >> 4) | ((uint64_t)(4096u >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61); }
__host__ __device__ constexpr uint32_t idesc(int m, int n, int amn) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amn << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t i, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(i), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t i, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(i), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}\n" ::"r"(s32(bar)), "r"(parity) : "memory");
}

struct Group
{
    int akind; // 0 K-major smem, 1 MN-major smem, 2 TMEM
    int n;     // 32 / 64 / 128
    int aoff;  // smem byte offset of the A tile, or TMEM column of A
    int dcol;  // accumulator column (within the stage's TMEM block)
    int acc0;  // accumulate flag of the first K-step
};
struct Stream
{
    Group g[8];
    int ng, per, ncommit;
};

__global__ void __launch_bounds__(128) tile_kernel(int iters, Stream st, unsigned long long *clk)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bar       = reinterpret_cast<uint64_t *>(smem + 98304);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 98304 + 64);
    for (int t = threadIdx.x; t < 98304 / 4; t += 128)
        reinterpret_cast<float *>(smem)[t] = 0.001f * (t % 97);
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0)
    {
        const uint32_t sb = s32(smem);
        const uint64_t b  = desc_k(sb + 65536);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
        {
            const int s       = it & 1;
            const uint32_t ts = tmem + s * 192;
            const uint32_t ab = sb + s * 32768;
            for (int gi = 0; gi < st.ng; ++gi)
            {
                const Group g     = st.g[gi];
                const uint32_t id = idesc(128, g.n, g.akind == 1);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    if (g.akind == 2)
                        mma_ts(ts + g.dcol, ts + g.aoff + 8 * k, b + 2 * k, id, (g.acc0 | k) > 0);
                    else if (g.akind == 1)
                        mma_ss(ts + g.dcol, desc_mn(ab + g.aoff) + 64 * k, b + 2 * k, id, (g.acc0 | k) > 0);
                    else
                        mma_ss(ts + g.dcol, desc_k(ab + g.aoff) + 2 * k, b + 2 * k, id, (g.acc0 | k) > 0);
                }
                if ((gi + 1) % st.per == 0)
                    for (int c = 0; c < st.ncommit; ++c)
                        commit(bar + 1 + c);
            }
        }
        commit(bar);
        mbar_wait(bar, 0);
        if (blockIdx.x == 0)
            clk[0] = (unsigned long long)(clock64() - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main()
{
    const size_t smem = 98304 + 128;
    cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned long long *clk, h;
    cudaMalloc(&clk, 8);
    struct Named { const char *what; Stream s; };
    const Named cases[] = {
        {"1 group  K-smem N32", {{{0, 32, 0, 64, 0}}, 1, 1, 0}},
        {"1 group  K-smem N64", {{{0, 64, 0, 64, 0}}, 1, 1, 0}},
        {"1 group  MN-smem N64", {{{1, 64, 0, 64, 0}}, 1, 1, 0}},
        {"1 group  TMEM N64", {{{2, 64, 0, 64, 0}}, 1, 1, 0}},
        {"2 groups K-smem N64 + N32 same D (dir0 SS)", {{{0, 64, 0, 64, 0}, {0, 32, 16384, 64, 1}}, 2, 2, 0}},
        {"2 groups K-smem N64 + N64 same D", {{{0, 64, 0, 64, 0}, {0, 64, 16384, 64, 1}}, 2, 2, 0}},
        {"2 groups K-smem N64 + N64 same D same A", {{{0, 64, 0, 64, 0}, {0, 64, 0, 64, 1}}, 2, 2, 0}},
        {"2 groups K-smem N64, different D", {{{0, 64, 0, 64, 0}, {0, 64, 16384, 128, 0}}, 2, 2, 0}},
        {"2 groups MN-smem N64 + N32 same D (dir1)", {{{1, 64, 0, 128, 0}, {1, 32, 16384, 128, 1}}, 2, 2, 0}},
        {"2 groups TMEM N64 + N32 same D (dir0 TS)", {{{2, 64, 0, 64, 0}, {2, 32, 32, 64, 1}}, 2, 2, 0}},
        {"2 groups TMEM N64 + N32, 2 commits", {{{2, 64, 0, 64, 0}, {2, 32, 32, 64, 1}}, 2, 2, 2}},
        {"tile: TS64 TS32 | MN64 MN32, 2+2 commits", {{{2, 64, 0, 64, 0}, {2, 32, 32, 64, 1}, {1, 64, 0, 128, 0}, {1, 32, 16384, 128, 1}}, 4, 2, 2}},
        {"tile: TS64 TS32 | MN64 MN32, no commits", {{{2, 64, 0, 64, 0}, {2, 32, 32, 64, 1}, {1, 64, 0, 128, 0}, {1, 32, 16384, 128, 1}}, 4, 2, 0}},
        {"tile interleaved: TS64 MN64 TS32 MN32", {{{2, 64, 0, 64, 0}, {1, 64, 0, 128, 0}, {2, 32, 32, 64, 1}, {1, 32, 16384, 128, 1}}, 4, 4, 0}},
        {"tile: 1 x TS N128(hi|lo|hi.. K=64 form) + MN64 MN32", {{{2, 128, 0, 64, 0}, {1, 64, 0, 192, 0}, {1, 32, 16384, 192, 1}}, 3, 3, 0}},
    };
    for (const Named &c : cases)
    {
        const int iters = 4000;
        tile_kernel<<<148, 128, smem>>>(100, c.s, clk);
        tile_kernel<<<148, 128, smem>>>(iters, c.s, clk);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
        printf("%-56s %8.1f clk per iteration = %6.1f per instruction   %s\n", c.what, (double)h / iters,
               (double)h / iters / (4 * c.s.ng), cudaGetErrorString(e));
    }
    return 0;
}

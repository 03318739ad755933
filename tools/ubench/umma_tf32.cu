// umma_tf32.cu -- tcgen05.mma kind::tf32 from shared memory on B200: (1) one 128 x N x 32 product checked against the
// host (validates the shared-memory / instruction descriptor encodings used by csrc/sumfac_umma.cuh), (2) issue rate of
// M = 128, K = 8 instructions for N = 32 ... 256 with both operands in 128B-swizzled K-major tiles.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_tf32 umma_tf32.cu && ./umma_tf32
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <vector>

#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// K-major, 128-byte swizzle: rows of 32 tf32 (128 B), atoms of 8 rows (1024 B, 1024-B aligned), 16-byte chunk c of row r
// sits at chunk position c ^ (r % 8).  SBO = 1024 B between 8-row groups; LBO unused (one atom along K).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);       // start address, bits [0,14)
    d |= (uint64_t)0 << 16;                         // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;              // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version 1 (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B, bits [61,64)
    return d;
}

// MN-major tf32 (A operand of the second contraction).  The only layout the tensor core accepts for 4-byte MN-major
// operands is SWIZZLE_128B_BASE32B (layout type 1; CUTLASS: "for mn-major tf32 operands, SW128_32B is the only available
// smem layout", Swizzle<2,5,2> on the byte address): 32 consecutive m (128 B) per k-row, atoms of 4 k-rows (512 B), the
// 32-byte chunk j of k-row r sits at chunk position j ^ (r % 4).  LBO = byte distance between 32-wide m blocks, SBO =
// between 4-deep k groups; one K = 8 instruction spans two k groups.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61; // SWIZZLE_128B_BASE32B
    return d;
}

// instruction descriptor: D = F32, A = B = TF32, dense; a_mn = 1: A is MN-major (else K-major); B K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn = 0)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                 "@P1 bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// element (row, k) of a K-major SW128 tile whose rows hold 32 floats
__device__ __host__ inline int sw128_index(int row, int k)
{
    return (row / 8) * 256 + (row % 8) * 32 + (((k / 4) ^ (row % 8)) * 4) + (k % 4);
}

// element (m, k) of an MN-major SW128_32B tile of 128 m x 32 k floats: m blocks of 32 at 1024 floats (LBO = 4096 B),
// k groups of 4 at 128 floats (SBO = 512 B)
__device__ __host__ inline int mn_sw128_index(int m, int k)
{
    return (m / 32) * 1024 + (k / 4) * 128 + (k % 4) * 32 + ((((m % 32) / 8) ^ (k % 4)) * 8) + (m % 8);
}

// N: instruction N; mode 0: correctness (one 128 x N x 32 product -> D), mode 1: rate (iters x 12 instructions)
template <int N> __global__ void __launch_bounds__(128) umma_kernel(const float *A, const float *B, float *D, int iters, int mode)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *sA            = reinterpret_cast<float *>(smem_raw);                 // 128 x 32 floats = 16 KB
    float *sB            = reinterpret_cast<float *>(smem_raw + 16384);         // N x 32 floats (<= 32 KB)
    uint64_t *bar        = reinterpret_cast<uint64_t *>(smem_raw + 16384 + 32768);
    uint32_t *tmem_slot  = reinterpret_cast<uint32_t *>(smem_raw + 16384 + 32768 + 8);
    const int tid = threadIdx.x, warp = tid >> 5;

    const bool a_mn = mode >= 2; // modes 2 / 3: the same checks with A stored MN-major
    for (int t = tid; t < 128 * 32; t += 128)
        sA[a_mn ? mn_sw128_index(t / 32, t % 32) : sw128_index(t / 32, t % 32)] = A[t];
    for (int t = tid; t < N * 32; t += 128)
        sB[sw128_index(t / 32, t % 32)] = B[t];
    if (tid == 0)
    {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0)
    {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = make_idesc_tf32(128, N, a_mn ? 1 : 0);
    const uint64_t adesc = a_mn ? make_desc_mn_sw128(smem_u32(sA), 4096u, 512u) : make_desc_sw128(smem_u32(sA));
    const uint64_t bdesc = make_desc_sw128(smem_u32(sB));
    const uint64_t astep = a_mn ? (1024u >> 4) : 2u; // next K = 8: two 4-deep k groups (MN-major) / 32 bytes (K-major)

    if (mode == 0 || mode == 2)
    {
        if (tid == 0)
        {
#pragma unroll
            for (int k = 0; k < 4; ++k) // K = 8 tf32 = 32 bytes per instruction: advance the start address by 2 (x16 B)
                umma_tf32(tmem, adesc + astep * k, bdesc + 2 * k, idesc, k > 0);
            umma_commit(bar);
        }
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int c0 = 0; c0 < N; c0 += 32)
        {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v); // warp w reads TMEM lanes 32w .. 32w+31
            for (int c = 0; c < 32; ++c)
                D[(size_t)tid * N + c0 + c] = __uint_as_float(v[c]);
        }
    }
    else
    {
        if (tid == 0)
        {
            for (int it = 0; it < iters; ++it)
            {
                const uint32_t acc = tmem + (uint32_t)((it & 1) * N); // two accumulators alternate
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_tf32(acc, adesc + astep * k, bdesc + 2 * k, idesc, (g | k) > 0);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (D && tid == 0 && blockIdx.x == 0)
            D[0] = 1.0f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

static float tf32_of(float x) // keep 10 mantissa bits
{
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

template <int N> void run(const float *dA, const float *dB, float *dD, const std::vector<float> &hA, const std::vector<float> &hB)
{
    const size_t smem = 16384 + 32768 + 64;
    cudaFuncSetAttribute(umma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int a_mn = 0; a_mn < 2; ++a_mn)
  {
    cudaMemset(dD, 0, 128 * 256 * 4);
    umma_kernel<N><<<1, 128, smem>>>(dA, dB, dD, 0, a_mn ? 2 : 0);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> hD(128 * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0.0, scale = 0.0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n)
        {
            double s = 0.0;
            for (int k = 0; k < 32; ++k)
                s += (double)hA[m * 32 + k] * (double)hB[n * 32 + k];
            worst = std::max(worst, std::abs(s - (double)hD[m * N + n]));
            scale = std::max(scale, std::abs(s));
        }
    printf("N=%3d A %s-major check: max|err| = %.3e (max|D| = %.3f)  %s\n", N, a_mn ? "MN" : "K", worst, scale,
           cudaGetErrorString(e));
    if (e != cudaSuccess)
        exit(1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int grid : {148, 296})
    {
        umma_kernel<N><<<grid, 128, smem>>>(dA, dB, nullptr, 100, a_mn ? 3 : 1);
        cudaEventRecord(e0);
        umma_kernel<N><<<grid, 128, smem>>>(dA, dB, nullptr, iters, a_mn ? 3 : 1);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 128 * N * 8 * 12.0 * iters * grid;
        printf("N=%3d A %s-major grid=%3d: %8.3f ms, %7.1f TFLOP/s tf32 (%.1f clk per M128 x N%d x K8 instruction at 1.965 GHz)  %s\n", N,
               a_mn ? "MN" : "K", grid, ms, flop / ms * 1e-9, ms * 1e-3 * 1.965e9 / (12.0 * iters * (grid / 148)), N, cudaGetErrorString(e));
    }
  }
}

int main()
{
    std::vector<float> hA(128 * 32), hB(256 * 32);
    srand(7);
    for (auto &v : hA)
        v = tf32_of((float)rand() / RAND_MAX - 0.5f);
    for (auto &v : hB)
        v = tf32_of((float)rand() / RAND_MAX - 0.5f);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, hA.size() * 4);
    cudaMalloc(&dB, hB.size() * 4);
    cudaMalloc(&dD, 128 * 256 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    run<32>(dA, dB, dD, hA, hB);
    run<64>(dA, dB, dD, hA, hB);
    run<128>(dA, dB, dD, hA, hB);
    run<256>(dA, dB, dD, hA, hB);
    return 0;
}

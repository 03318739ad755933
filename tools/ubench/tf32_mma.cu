// tf32_mma.cu -- rate of the warp-level TF32 tensor-core path (mma.sync.m16n8k8.tf32, FP32 accumulate)
// on sm_100a: is a 3xTF32 split (FP32-accurate products) faster than plain FFMA (~72 TFLOP/s)?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32_mma tf32_mma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CH> __global__ void k(float *out, int iters)
{
    float c[CH][4];
    unsigned a[4], b[2];
    for (int i = 0; i < CH; ++i)
        for (int j = 0; j < 4; ++j)
            c[i][j] = threadIdx.x * 1e-3f + i + j;
    for (int j = 0; j < 4; ++j)
        a[j] = __float_as_uint(1.0f + threadIdx.x * 1e-4f + j * 1e-3f) & 0xffffe000u;
    for (int j = 0; j < 2; ++j)
        b[j] = __float_as_uint(1.0f - threadIdx.x * 1e-4f - j * 1e-3f) & 0xffffe000u;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int s = 0; s < 16; ++s)
#pragma unroll
            for (int i = 0; i < CH; ++i)
                mma_tf32(c[i], a, b);
    float r = 0;
    for (int i = 0; i < CH; ++i)
        for (int j = 0; j < 4; ++j)
            r += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int CH> void run(int warps_per_sm)
{
    float *out;
    const int threads = 128, blocks = 148 * warps_per_sm * 32 / threads, iters = 400;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(float));
    k<CH><<<blocks, threads>>>(out, iters);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<CH><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = (double)blocks * threads / 32 * iters * 16.0 * CH * 2.0 * 16 * 8 * 8;
    std::printf("tf32 m16n8k8 chains=%d warps/SM=%2d  %8.2f TFLOP/s (%.3f ms)  -> 3xTF32 equivalent %7.2f  %s\n", CH,
                warps_per_sm, fl / ms * 1e-9, ms, fl / ms * 1e-9 / 3, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    for (int w : {4, 8, 16, 32})
    {
        run<1>(w);
        run<2>(w);
        run<4>(w);
        run<8>(w);
    }
    return 0;
}

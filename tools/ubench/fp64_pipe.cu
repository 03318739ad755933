// fp64_pipe.cu -- micro-benchmarks of the FP64/FP32 FMA pipe fed from the constant bank, to size the
// sum-factorisation kernels: how close to peak can DFMA/FFMA run when one operand is a uniform
// constant-bank value, as a function of chains per thread, constant-load width and resident warps?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

static __constant__ __align__(16) double cbd[2048];
static __constant__ __align__(16) float cbf[2048];

// MODE 0: operand = register (no loads)          -> pure pipe rate
// MODE 1: immediate constant slots, unrolled       (ptxas picks LDCU.64/.128)
// MODE 2: runtime-indexed constant slots (LDCU.64 c[3][UR])
template <typename T, int CH, int MODE> __global__ void fma_kernel(T *out, int iters, int zero)
{
    T acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c)
        acc[c] = (T)threadIdx.x * (T)1e-3 + (T)c;
    const T x = (T)1.0000001 + (T)threadIdx.x * (T)1e-9;
    const T *cb = sizeof(T) == 8 ? (const T *)cbd : (const T *)cbf;
    for (int it = 0; it < iters; ++it)
    {
        if (MODE == 0)
        {
#pragma unroll
            for (int s = 0; s < 64; ++s)
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    acc[c] = fma(acc[c], x, x);
        }
        else if (MODE == 1)
        {
#pragma unroll
            for (int s = 0; s < 64; ++s)
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    acc[c] = fma(x, cb[(s * CH + c) % 1024], acc[c]);
        }
        else
        {
            const int base = (it * zero) & 1023;
#pragma unroll 1
            for (int s = 0; s < 64; ++s)
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    acc[c] = fma(x, cb[base + s * CH + c], acc[c]);
        }
    }
    T r = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c)
        r += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename T, int CH, int MODE> void run(const char *name, int warps_per_sm)
{
    T *out;
    const int threads = 128;
    const int blocks  = 148 * warps_per_sm * 32 / threads;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(T));
    const int iters = 200;
    fma_kernel<T, CH, MODE><<<blocks, threads>>>(out, iters, 0);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    fma_kernel<T, CH, MODE><<<blocks, threads>>>(out, iters, 0);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double flops = 2.0 * (double)blocks * threads * iters * 64 * CH;
    std::printf("%-4s chains=%2d mode=%d warps/SM=%2d  %8.2f TFLOP/s  (%.3f ms)\n", name, CH, MODE, warps_per_sm,
                flops / ms * 1e-9, ms);
    cudaFree(out);
}

int main()
{
    for (int w : {4, 8, 16, 32})
    {
        run<double, 1, 0>("f64", w);
        run<double, 2, 0>("f64", w);
        run<double, 4, 0>("f64", w);
        run<double, 8, 0>("f64", w);
        run<double, 16, 0>("f64", w);
        run<double, 4, 1>("f64", w);
        run<double, 8, 1>("f64", w);
        run<double, 16, 1>("f64", w);
        run<double, 4, 2>("f64", w);
        run<double, 8, 2>("f64", w);
        run<double, 16, 2>("f64", w);
    }
    for (int w : {8, 16, 32})
    {
        run<float, 4, 0>("f32", w);
        run<float, 8, 0>("f32", w);
        run<float, 16, 0>("f32", w);
        run<float, 8, 1>("f32", w);
        run<float, 16, 1>("f32", w);
        run<float, 8, 2>("f32", w);
        run<float, 16, 2>("f32", w);
    }
    return 0;
}

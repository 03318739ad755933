// dmma_pipe.cu -- does sm_100a run FP64 tensor-core MMA (mma.sync ... f64) at a useful rate, and does it
// share the DFMA pipe?  Measures m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 with CH independent accumulators,
// alone and interleaved 1:1 (in FLOPs) with plain DFMA chains.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_pipe dmma_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE> struct Frag; // registers per lane: A, B, C
template <> struct Frag<884>  { static constexpr int A = 1, B = 1, C = 2; static constexpr double FL = 2.0 * 8 * 8 * 4; };
template <> struct Frag<1684> { static constexpr int A = 2, B = 1, C = 4; static constexpr double FL = 2.0 * 16 * 8 * 4; };
template <> struct Frag<1688> { static constexpr int A = 4, B = 2, C = 4; static constexpr double FL = 2.0 * 16 * 8 * 8; };
template <> struct Frag<16816>{ static constexpr int A = 8, B = 4, C = 4; static constexpr double FL = 2.0 * 16 * 8 * 16; };

template <int SHAPE> __device__ __forceinline__ void mma(double *c, const double *a, const double *b)
{
    if constexpr (SHAPE == 884)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
    else if constexpr (SHAPE == 1684)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
    else if constexpr (SHAPE == 1688)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                     "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
                       "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// FMIX: number of DFMA chains run alongside (0 = MMA only); each chain does `FPER` DFMAs per MMA round
template <int SHAPE, int CH, int FMIX, int FPER> __global__ void k(double *out, int iters)
{
    using F = Frag<SHAPE>;
    double c[CH][F::C], a[F::A], b[F::B], f[FMIX > 0 ? FMIX : 1];
    for (int i = 0; i < CH; ++i)
        for (int j = 0; j < F::C; ++j)
            c[i][j] = threadIdx.x * 1e-3 + i + j;
    for (int j = 0; j < F::A; ++j)
        a[j] = 1.0 + threadIdx.x * 1e-9 + j * 1e-7;
    for (int j = 0; j < F::B; ++j)
        b[j] = 1.0 - threadIdx.x * 1e-9 - j * 1e-7;
    for (int j = 0; j < FMIX; ++j)
        f[j] = threadIdx.x + j;
    const double x = 1.0000001 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int s = 0; s < 16; ++s)
        {
#pragma unroll
            for (int i = 0; i < CH; ++i)
                mma<SHAPE>(c[i], a, b);
#pragma unroll
            for (int r = 0; r < FPER; ++r)
#pragma unroll
                for (int j = 0; j < FMIX; ++j)
                    f[j] = fma(f[j], x, x);
        }
    }
    double r = 0;
    for (int i = 0; i < CH; ++i)
        for (int j = 0; j < F::C; ++j)
            r += c[i][j];
    for (int j = 0; j < FMIX; ++j)
        r += f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int SHAPE, int CH, int FMIX, int FPER> void run(int warps_per_sm)
{
    double *out;
    const int threads = 128, blocks = 148 * warps_per_sm * 32 / threads, iters = 200;
    cudaMalloc(&out, (size_t)blocks * threads * sizeof(double));
    k<SHAPE, CH, FMIX, FPER><<<blocks, threads>>>(out, iters);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<SHAPE, CH, FMIX, FPER><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * threads / 32;
    const double mma_fl = warps * iters * 16.0 * CH * Frag<SHAPE>::FL;
    const double fma_fl = warps * 32 * iters * 16.0 * FMIX * FPER * 2.0;
    std::printf("m%-6d ch=%d fma_chains=%d x%d warps/SM=%2d  mma %7.2f TF/s  dfma %7.2f TF/s  total %7.2f  (%.3f ms) %s\n",
                SHAPE, CH, FMIX, FPER, warps_per_sm, mma_fl / ms * 1e-9, fma_fl / ms * 1e-9, (mma_fl + fma_fl) / ms * 1e-9, ms,
                cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    for (int w : {4, 8, 16, 32})
    {
        run<884, 1, 0, 0>(w);
        run<884, 2, 0, 0>(w);
        run<884, 4, 0, 0>(w);
        run<884, 8, 0, 0>(w);
        run<1684, 4, 0, 0>(w);
        run<1688, 4, 0, 0>(w);
        run<16816, 2, 0, 0>(w);
        run<16816, 4, 0, 0>(w);
    }
    // shared pipe?  4 MMA chains (m8n8k4: 8 FMA/lane each) + 8 DFMA chains x 4 = same FLOPs both sides
    for (int w : {8, 16, 32})
    {
        run<884, 4, 8, 4>(w);
        run<884, 4, 8, 1>(w);
        run<16816, 4, 8, 4>(w);
    }
    return 0;
}

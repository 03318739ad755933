// stream_add.cu -- which launch shape streams x += y fastest on a B200?  (benchmark02's kernel; the reference's own
// add_vector<T,true><<<n/8/1024, 1024>>> measured 7.1 TB/s on this pool, above the library's first version.)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stream_add stream_add.cu && ./stream_add
// Variants: THREADS per CTA, U independent 16-byte vectors per thread and array, persistent grid-stride (P = 1) or one
// chunk per CTA (P = 0); ld/st cache hints streaming (.cs) or default.
#include <cstdio>
#include <cuda_runtime.h>

template <int THREADS, int U, bool PERSIST, bool CS>
__global__ void __launch_bounds__(THREADS) add_kernel(double2 *__restrict__ x, const double2 *__restrict__ y, size_t nv)
{
    const size_t chunk = (size_t)THREADS * U;
    for (size_t c = blockIdx.x; c * chunk < nv; c += gridDim.x)
    {
        const size_t base = c * chunk + threadIdx.x;
        double2 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (base + (size_t)u * THREADS < nv)
            {
                a[u] = CS ? __ldcs(x + base + (size_t)u * THREADS) : x[base + (size_t)u * THREADS];
                b[u] = CS ? __ldcs(y + base + (size_t)u * THREADS) : __ldg(y + base + (size_t)u * THREADS);
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (base + (size_t)u * THREADS < nv)
            {
                a[u].x += b[u].x;
                a[u].y += b[u].y;
                if (CS)
                    __stcs(x + base + (size_t)u * THREADS, a[u]);
                else
                    x[base + (size_t)u * THREADS] = a[u];
            }
        if (!PERSIST)
            break;
    }
}

template <int THREADS, int U, bool PERSIST, bool CS> void run(double2 *x, double2 *y, size_t nv, int mult)
{
    int sms = 148, occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, add_kernel<THREADS, U, PERSIST, CS>, THREADS, 0);
    const size_t chunks = (nv + (size_t)THREADS * U - 1) / ((size_t)THREADS * U);
    const unsigned grid = PERSIST ? (unsigned)(sms * occ * mult) : (unsigned)chunks;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 8; ++r)
    {
        cudaEventRecord(e0);
        add_kernel<THREADS, U, PERSIST, CS><<<grid, THREADS>>>(x, y, nv);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best)
            best = ms;
    }
    printf("threads=%4d U=%d persist=%d(x%d) cs=%d occ=%d grid=%8u  %8.4f ms  %7.1f GB/s  %s\n", THREADS, U, (int)PERSIST,
           mult, (int)CS, occ, grid, best, 48.0 * nv / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    for (size_t lg : {28, 30})
    {
        const size_t n = (size_t)1 << lg, nv = n / 2;
        double2 *x, *y;
        cudaMalloc(&x, n * 8);
        cudaMalloc(&y, n * 8);
        cudaMemset(x, 0, n * 8);
        cudaMemset(y, 0, n * 8);
        printf("---- n = 2^%zu doubles\n", lg);
        run<256, 2, true, true>(x, y, nv, 1);
        run<256, 4, true, true>(x, y, nv, 1);
        run<256, 4, true, true>(x, y, nv, 4);
        run<256, 1, false, true>(x, y, nv, 1);
        run<256, 2, false, true>(x, y, nv, 1);
        run<256, 4, false, true>(x, y, nv, 1);
        run<256, 8, false, true>(x, y, nv, 1);
        run<512, 4, false, true>(x, y, nv, 1);
        run<1024, 1, false, true>(x, y, nv, 1);
        run<1024, 2, false, true>(x, y, nv, 1);
        run<1024, 4, false, true>(x, y, nv, 1);
        run<1024, 4, false, false>(x, y, nv, 1);
        run<256, 4, false, false>(x, y, nv, 1);
        run<1024, 4, true, true>(x, y, nv, 1);
        run<1024, 4, true, false>(x, y, nv, 1);
        run<128, 4, false, true>(x, y, nv, 1);
        run<128, 8, false, true>(x, y, nv, 1);
        cudaFree(x);
        cudaFree(y);
    }
    return 0;
}

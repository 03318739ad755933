// iprod_probe.cu -- timing of the tile shapes of the lanes-style IProductWRTBase kernels (csrc/sumfac_iprod_lanes.cuh)
// at ~64 Mi quadrature points, unweighted and weighted.  Correctness is the tests' job (tests/test_iproduct_gpu.py
// forces every routed shape); here every variant of a case must at least produce the same checksum.
#include <cstdio>
#include <cstdlib>

#include "../../gpu-benchmarking_b200/csrc/sumfac_iprod_lanes.cuh"

using namespace b200fe;
namespace b200fe
{
std::atomic<unsigned long long> g_launch_count{0};
thread_local const char *t_last_backend = "";
}

#define CK(x)                                                                                                \
    do                                                                                                       \
    {                                                                                                        \
        cudaError_t e_ = (x);                                                                                \
        if (e_ != cudaSuccess)                                                                               \
        {                                                                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);                  \
            exit(1);                                                                                         \
        }                                                                                                    \
    } while (0)

template <typename T> __global__ void fill_kernel(T *x, size_t n, unsigned seed)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        unsigned h = (unsigned)i * 2654435761u + seed;
        h ^= h >> 15;
        h *= 2246822519u;
        h ^= h >> 13;
        x[i] = (T)((double)(h & 0xffffff) / 16777216.0 + 0.25);
    }
}
template <typename T> __global__ void xor_kernel(const T *a, size_t n, unsigned long long *acc)
{
    unsigned long long c = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        c ^= (sizeof(T) == 8 ? reinterpret_cast<const unsigned long long *>(a)[i]
                             : (unsigned long long)reinterpret_cast<const unsigned *>(a)[i]) * (2 * i + 1);
    atomicXor(acc, c);
}

template <typename T> struct Case
{
    int dim, nq;
    unsigned nelmt;
    T *b[3], *in, *w, *out;
    size_t nin, nout;
    unsigned long long *acc;
    BankGuard bank;
    void setup(int dim_, int nq_)
    {
        dim = dim_;
        nq  = nq_;
        const int nm = nq - 1;
        size_t pts = 1, modes = 1;
        for (int d = 0; d < dim; ++d)
            pts *= nq, modes *= nm;
        nelmt = (unsigned)((64ull << 20) / pts) - 3;
        nin   = pts * nelmt;
        nout  = modes * nelmt;
        for (int d = 0; d < 3; ++d)
        {
            CK(cudaMalloc(&b[d], sizeof(T) * nm * nq));
            fill_kernel<T><<<1, 256>>>(b[d], (size_t)nm * nq, 17u + d);
        }
        CK(cudaMalloc(&in, sizeof(T) * nin));
        CK(cudaMalloc(&w, sizeof(T) * nin));
        CK(cudaMalloc(&out, sizeof(T) * nout));
        CK(cudaMalloc(&acc, 8));
        fill_kernel<T><<<1024, 256>>>(in, nin, 99u);
        fill_kernel<T><<<1024, 256>>>(w, nin, 7u);
        const T *bs[3] = {b[0], b[1], b[2]};
        if (fill_basis_bank<T>(bank, dim, bs, nm, nq, true, 0))
        {
            printf("bank fill failed\n");
            exit(1);
        }
        CK(cudaDeviceSynchronize());
    }
    void teardown()
    {
        for (int d = 0; d < 3; ++d)
            cudaFree(b[d]);
        cudaFree(in);
        cudaFree(w);
        cudaFree(out);
        cudaFree(acc);
    }
    template <typename K> void run(const char *name, int weighted, K kernel, unsigned grid, int threads, size_t smem)
    {
        if (smem > 227 * 1024)
            return;
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
        CK(cudaMemset(out, 0xff, sizeof(T) * nout));
        CK(cudaMemset(acc, 0, 8));
        kernel<<<grid, threads, smem>>>(in, weighted ? w : nullptr, out, nelmt);
        CK(cudaGetLastError());
        xor_kernel<T><<<1024, 256>>>(out, nout, acc);
        unsigned long long h = 0;
        CK(cudaMemcpy(&h, acc, 8, cudaMemcpyDeviceToHost));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float sum = 0;
        const int reps = 12;
        for (int r = 0; r < reps; ++r)
        {
            cudaEventRecord(e0);
            kernel<<<grid, threads, smem>>>(in, weighted ? w : nullptr, out, nelmt);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 2)
                sum += ms;
        }
        const double bytes = (double)sizeof(T) * ((1 + weighted) * nin + nout);
        printf("%s,%d,%s,%d,%s,%d,%zu,%d,%.4f,%.3f,%016llx\n", dim == 2 ? "quad" : "hex", nq, sizeof(T) == 8 ? "f64" : "f32",
               weighted, name, threads, smem, occ, sum / (reps - 2), bytes / (sum / (reps - 2)) * 1e-6 / 6546.9, h);
        fflush(stdout);
    }
};

#define IQ(T, NQ, EL)                                                                                        \
    c.run("EL=" #EL, 0, iproduct_quad_lanes_kernel<T, NQ, EL, false>, (c.nelmt + EL - 1) / EL,               \
          QuadIprodLanes<T, NQ, EL>::THREADS, QuadIprodLanes<T, NQ, EL>::SMEM);                               \
    c.run("EL=" #EL, 1, iproduct_quad_lanes_kernel<T, NQ, EL, true>, (c.nelmt + EL - 1) / EL,                \
          QuadIprodLanes<T, NQ, EL>::THREADS, QuadIprodLanes<T, NQ, EL>::SMEM);
#define IH(T, NQ, EL, MB)                                                                                    \
    c.run("EL=" #EL "/MINB=" #MB, 0, iproduct_hex_lanes_kernel<T, NQ, EL, false, MB>, (c.nelmt + EL - 1) / EL, \
          HexIprodLanes<T, NQ, EL>::THREADS, HexIprodLanes<T, NQ, EL>::SMEM);                                 \
    c.run("EL=" #EL "/MINB=" #MB, 1, iproduct_hex_lanes_kernel<T, NQ, EL, true, MB>, (c.nelmt + EL - 1) / EL, \
          HexIprodLanes<T, NQ, EL>::THREADS, HexIprodLanes<T, NQ, EL>::SMEM);

#define IS(T, NQ, EL, MB)                                                                                    \
    c.run("staged EL=" #EL "/MINB=" #MB, 0, iproduct_hex_lanes_kernel<T, NQ, EL, false, MB, true>,            \
          (c.nelmt + EL - 1) / EL, HexIprodLanes<T, NQ, EL, true>::THREADS, HexIprodLanes<T, NQ, EL, true>::SMEM); \
    c.run("staged EL=" #EL "/MINB=" #MB, 1, iproduct_hex_lanes_kernel<T, NQ, EL, true, MB, true>,             \
          (c.nelmt + EL - 1) / EL, HexIprodLanes<T, NQ, EL, true>::THREADS, HexIprodLanes<T, NQ, EL, true>::SMEM);

int main()
{
    printf("op,nq,dtype,weighted,shape,threads,smem,ctas_per_sm,ms_avg,hbm_frac,checksum\n");
    {
        Case<double> c;
        c.setup(3, 4); IH(double, 4, 32, 1) IS(double, 4, 32, 1) IS(double, 4, 64, 1) IS(double, 4, 16, 1) c.teardown();
        c.setup(3, 6); IH(double, 6, 4, 1) IS(double, 6, 4, 1) IS(double, 6, 8, 1) IS(double, 6, 16, 1) c.teardown();
        c.setup(3, 8); IH(double, 8, 4, 1) IS(double, 8, 4, 1) IS(double, 8, 8, 1) IS(double, 8, 16, 1) c.teardown();
    }
    {
        Case<float> c;
        c.setup(3, 4); IH(float, 4, 64, 1) IS(float, 4, 64, 1) IS(float, 4, 32, 1) c.teardown();
        c.setup(3, 6); IH(float, 6, 8, 1) IS(float, 6, 8, 1) IS(float, 6, 16, 1) IS(float, 6, 32, 1) c.teardown();
        c.setup(3, 8); IH(float, 8, 4, 1) IS(float, 8, 4, 1) IS(float, 8, 8, 1) IS(float, 8, 16, 1) IS(float, 8, 32, 1) c.teardown();
        c.setup(3, 10); IH(float, 10, 8, 3) IS(float, 10, 8, 1) IS(float, 10, 8, 3) IS(float, 10, 16, 1) IS(float, 10, 16, 2) c.teardown();
    }
    return 0;
}

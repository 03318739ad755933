// lanesem_probe.cu -- on-device comparison of the tile shapes of the element-major "lanes-em" quad kernel
// (csrc/sumfac_lanes.cuh) at ~64 Mi quadrature points, checked bit for bit against the generic kernel.
#include <cstdio>
#include <cstdlib>

#include "../../gpu-benchmarking_b200/csrc/sumfac_generic.cuh"
#include "../../gpu-benchmarking_b200/csrc/sumfac_lanes.cuh"

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
        x[i] = (T)((double)(h & 0xffffff) / 16777216.0 - 0.5);
    }
}
template <typename T> __global__ void diff_kernel(const T *a, const T *b, size_t n, unsigned long long *bad)
{
    unsigned long long c = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        if (sizeof(T) == 8)
            c += (reinterpret_cast<const unsigned long long *>(a)[i] != reinterpret_cast<const unsigned long long *>(b)[i]);
        else
            c += (reinterpret_cast<const unsigned *>(a)[i] != reinterpret_cast<const unsigned *>(b)[i]);
    }
    if (c)
        atomicAdd(bad, c);
}

template <typename T> struct Case
{
    int dim = 2;
    int nq;
    unsigned nelmt;
    T *b[3], *in, *out, *ref;
    size_t nin, nout;
    unsigned long long *bad;
    BankGuard bank;

    void setup(int nq_, int dim_ = 2)
    {
        nq  = nq_;
        dim = dim_;
        const int nm = nq - 1;
        size_t pts = 1, modes = 1;
        for (int d = 0; d < dim; ++d)
            pts *= nq, modes *= nm;
        nelmt = (unsigned)((64ull << 20) / pts) - 3; // ragged last tile
        nin   = modes * nelmt;
        nout  = pts * nelmt;
        for (int d = 0; d < 3; ++d)
        {
            CK(cudaMalloc(&b[d], sizeof(T) * nm * nq));
            fill_kernel<T><<<1, 256>>>(b[d], (size_t)nm * nq, 17u + d);
        }
        CK(cudaMalloc(&in, sizeof(T) * nin));
        CK(cudaMalloc(&out, sizeof(T) * nout));
        CK(cudaMalloc(&ref, sizeof(T) * nout));
        CK(cudaMalloc(&bad, 8));
        fill_kernel<T><<<1024, 256>>>(in, nin, 99u);
        const unsigned n = nm, q = nq;
        if (dim == 2)
        {
            const size_t smem = sizeof(T) * (2 * n * q + n * n + q * n);
            bwdtrans_quad_generic_kernel<T><<<148 * 8, 128, smem>>>(n, n, q, q, nelmt, b[0], b[1], in, ref, 0);
        }
        else
        {
            const size_t smem = sizeof(T) * (3 * n * q + n * n * n + q * n * n + q * q * n);
            CK(cudaFuncSetAttribute(bwdtrans_hex_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            bwdtrans_hex_generic_kernel<T><<<148 * 4, 256, smem>>>(n, n, n, q, q, q, nelmt, b[0], b[1], b[2], in, ref, 0);
        }
        CK(cudaDeviceSynchronize());
        const T *bs[3] = {b[0], b[1], b[2]};
        if (fill_basis_bank<T>(bank, dim, bs, nm, nq, false, 0))
        {
            printf("bank fill failed\n");
            exit(1);
        }
        CK(cudaDeviceSynchronize());
    }
    void teardown()
    {
        cudaFree(b[0]);
        cudaFree(b[1]);
        cudaFree(b[2]);
        cudaFree(in);
        cudaFree(out);
        cudaFree(ref);
        cudaFree(bad);
    }
    template <typename K, typename... A>
    void run(const char *name, K kernel, unsigned grid, int threads, size_t smem, A... extra)
    {
        if (smem > 227 * 1024)
        {
            printf("# skipped %s (smem %zu)\n", name, smem);
            return;
        }
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
        CK(cudaMemset(out, 0xff, sizeof(T) * nout));
        CK(cudaMemset(bad, 0, 8));
        kernel<<<grid, threads, smem>>>(in, out, nelmt, extra...);
        CK(cudaGetLastError());
        diff_kernel<T><<<1024, 256>>>(out, ref, nout, bad);
        unsigned long long nbad = 0;
        CK(cudaMemcpy(&nbad, bad, 8, cudaMemcpyDeviceToHost));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e30f, sum = 0;
        const int reps = 12;
        for (int r = 0; r < reps; ++r)
        {
            cudaEventRecord(e0);
            kernel<<<grid, threads, smem>>>(in, out, nelmt, extra...);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
            if (r >= 2)
                sum += ms;
        }
        const double bytes = (double)sizeof(T) * (nin + nout);
        printf("%s,%d,%s,%s,%d,%zu,%d,%.4f,%.4f,%.1f,%.3f,%llu\n", dim == 2 ? "quad" : "hex", nq, sizeof(T) == 8 ? "f64" : "f32", name, threads, smem,
               occ, best, sum / (reps - 2), bytes / (sum / (reps - 2)) * 1e-6, bytes / (sum / (reps - 2)) * 1e-6 / 6546.9, nbad);
        fflush(stdout);
    }
};

#define QE(T, NQ, EL, MB)                                                                                    \
    c.run("EL=" #EL "/MINB=" #MB, bwdtrans_quad_lanesem_kernel<T, NQ, EL, MB>, (c.nelmt + EL - 1) / EL,       \
          QuadLanesEm<T, NQ, EL>::THREADS, QuadLanesEm<T, NQ, EL>::SMEM, (double *)nullptr);

#define HE(T, NQ, EL, MB)                                                                                    \
    c.run("EL=" #EL "/MINB=" #MB, bwdtrans_hex_lanesem_kernel<T, NQ, EL, MB>, (c.nelmt + EL - 1) / EL,        \
          HexLanesEm<T, NQ, EL>::THREADS, HexLanesEm<T, NQ, EL>::SMEM, (double *)nullptr);

#define QT(T, NQ, EL, TPC)                                                                                   \
    c.run("EL=" #EL "/TPC=" #TPC, bwdtrans_quad_lanesem_kernel<T, NQ, EL, 1, TPC>,                            \
          ((c.nelmt + EL - 1) / EL + TPC - 1) / TPC, QuadLanesEm<T, NQ, EL, TPC>::THREADS,                    \
          QuadLanesEm<T, NQ, EL, TPC>::SMEM, (double *)nullptr);
int main()
{
    printf("op,nq,dtype,shape,threads,smem,ctas_per_sm,ms_best,ms_avg,gb_s,hbm_frac,mismatches\n");
    {
        Case<float> c;
        c.setup(6); QT(float, 6, 16, 8) QT(float, 6, 12, 8) QT(float, 6, 24, 4) QT(float, 6, 20, 8) c.teardown();
        c.setup(10); QT(float, 10, 8, 4) QT(float, 10, 12, 4) QT(float, 10, 12, 2) QT(float, 10, 20, 2) c.teardown();
        c.setup(14); QT(float, 14, 8, 1) QT(float, 14, 12, 1) QT(float, 14, 4, 1) QT(float, 14, 12, 2) c.teardown();
        c.setup(16); QT(float, 16, 8, 1) QT(float, 16, 12, 1) QT(float, 16, 4, 1) c.teardown();
        c.setup(12); QT(float, 12, 16, 1) QT(float, 12, 12, 1) QT(float, 12, 20, 1) c.teardown();
    }
    return 0;
}

// lanes_probe.cu -- on-device comparison of the tile shapes of the interleaved-layout "lanes" kernels
// (csrc/sumfac_lanes.cuh): times every variant listed in main() at ~64 Mi quadrature points and checks it
// bit for bit against the run-time-size generic kernel.  Build: see tools/README.md.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../gpu-benchmarking_b200/csrc/sumfac_generic.cuh"
#include "../../gpu-benchmarking_b200/csrc/sumfac_lanes.cuh"
#include "../../gpu-benchmarking_b200/csrc/sumfac_coapipe.cuh"
#include "../../gpu-benchmarking_b200/csrc/sumfac_coamma.cuh"

using namespace b200fe;
namespace b200fe
{
std::atomic<unsigned long long> g_launch_count{0};
thread_local const char *t_last_backend = "";
thread_local unsigned long long t_bank_tag = 0;
std::atomic<int> g_bank_fill_mode{0};
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

template <typename T> __global__ void maxerr_kernel(const T *a, const T *b, size_t n, unsigned *err_bits, unsigned *ref_bits)
{
    float e = 0.f, m = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        e = fmaxf(e, fabsf((float)a[i] - (float)b[i]));
        m = fmaxf(m, fabsf((float)b[i]));
    }
    atomicMax(err_bits, __float_as_uint(e)); // non-negative floats order like their bit patterns
    atomicMax(ref_bits, __float_as_uint(m));
}

template <typename T> struct Case
{
    int dim, nq;
    unsigned nelmt;
    T *b[3], *in, *out, *ref;
    size_t nin, nout;
    unsigned long long *bad;
    BankGuard bank;

    void setup(int dim_, int nq_)
    {
        dim = dim_;
        nq  = nq_;
        const int nm = nq - 1;
        size_t pts = 1, modes = 1;
        for (int d = 0; d < dim; ++d)
            pts *= nq, modes *= nm;
        nelmt = (unsigned)((64ull << 20) / pts) / 32 * 32;
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
            bwdtrans_quad_generic_kernel<T><<<148 * 8, 128, smem>>>(n, n, q, q, nelmt, b[0], b[1], in, ref, 1);
        }
        else
        {
            const size_t smem = sizeof(T) * (3 * n * q + n * n * n + q * n * n + q * q * n);
            CK(cudaFuncSetAttribute(bwdtrans_hex_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            bwdtrans_hex_generic_kernel<T><<<148 * 4, 256, smem>>>(n, n, n, q, q, q, nelmt, b[0], b[1], b[2], in, ref, 1);
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
        for (int d = 0; d < 3; ++d)
            cudaFree(b[d]);
        cudaFree(in);
        cudaFree(out);
        cudaFree(ref);
        cudaFree(bad);
    }
    template <typename K> void run(const char *name, K kernel, unsigned grid, int threads, size_t smem)
    {
        run_args(name, kernel, grid, threads, smem, in, out, nelmt);
    }
    // any kernel signature: the arguments are passed through
    template <typename K, typename... A> void run_args(const char *name, K kernel, unsigned grid, int threads, size_t smem, A... args)
    {
        if (smem > 48 * 1024)
            CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
        if (grid == 0) // persistent kernels: one wave
            grid = 148 * (occ > 0 ? occ : 1);
        CK(cudaMemset(out, 0xff, sizeof(T) * nout));
        CK(cudaMemset(bad, 0, 8));
        kernel<<<grid, threads, smem>>>(args...);
        CK(cudaGetLastError());
        diff_kernel<T><<<1024, 256>>>(out, ref, nout, bad);
        unsigned long long nbad = 0;
        CK(cudaMemcpy(&nbad, bad, 8, cudaMemcpyDeviceToHost));
        if (nbad) // tensor-core FP32 kernels are not bit-identical: report the error relative to the largest output
        {
            unsigned *eb;
            CK(cudaMalloc(&eb, 8));
            CK(cudaMemset(eb, 0, 8));
            maxerr_kernel<T><<<1024, 256>>>(out, ref, nout, eb, eb + 1);
            unsigned h[2];
            CK(cudaMemcpy(h, eb, 8, cudaMemcpyDeviceToHost));
            float e, m;
            memcpy(&e, &h[0], 4);
            memcpy(&m, &h[1], 4);
            printf("#   max |diff| %.3e, max |ref| %.3e, ratio %.3e\n", e, m, e / m);
            cudaFree(eb);
        }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e30f, sum = 0;
        const int reps = 12;
        for (int r = 0; r < reps; ++r)
        {
            cudaEventRecord(e0);
            kernel<<<grid, threads, smem>>>(args...);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
            if (r >= 2)
                sum += ms;
        }
        const double bytes = (double)sizeof(T) * (nin + nout);
        printf("%s,%d,%s,%d,%d,%zu,%d,%.4f,%.4f,%.1f,%llu\n", dim == 2 ? "quad" : "hex", nq, sizeof(T) == 8 ? "f64" : "f32",
               0, threads, smem, occ, best, sum / (reps - 2), bytes / (sum / (reps - 2)) * 1e-6, nbad);
        printf("#   ^ %s\n", name);
        fflush(stdout);
    }
};

#define HP(T, NQ, EL, MB)                                                                                    \
    c.run("plane EL=" #EL " MINB=" #MB, bwdtrans_hex_lanes_kernel<T, NQ, EL, MB>, c.nelmt / EL,               \
          HexLanes<T, NQ, EL>::THREADS, HexLanes<T, NQ, EL>::SMEM);
#define HQ(T, NQ, EL, IH, MB)                                                                                \
    c.run("q-outer EL=" #EL " IH=" #IH " MINB=" #MB, bwdtrans_hex_lanesq_kernel<T, NQ, EL, IH, MB>, c.nelmt / EL, \
          HexLanesQ<T, NQ, EL, IH>::THREADS, HexLanesQ<T, NQ, EL, IH>::SMEM);
#define QL(T, NQ, EL)                                                                                        \
    c.run("quad EL=" #EL, bwdtrans_quad_lanes_kernel<T, NQ, EL>, c.nelmt / EL, QuadLanes<T, NQ, EL>::THREADS,  \
          QuadLanes<T, NQ, EL>::SMEM);


// coa-pipe (sumfac_coapipe.cuh): persistent (grid 0 = one wave of resident CTAs)
#define CP(T, NQ, EL, NW, MB)                                                                                \
    c.run_args("coa-pipe EL=" #EL " NW=" #NW " MINB=" #MB, bwdtrans_hex_coapipe_kernel<T, NQ, EL, NW, MB>, 0u,  \
               HexCoaPipe<T, NQ, EL, NW>::THREADS, HexCoaPipe<T, NQ, EL, NW>::SMEM, (const T *)c.in, c.out, c.nelmt / EL);
#define CPT(T, NQ, EL, NW, MB)                                                                               \
    {                                                                                                        \
        CUtensorMap tm;                                                                                      \
        if (!make_coa_tensor_map<T>(&tm, c.in, HexCoaPipe<T, NQ, EL, NW>::NM3, c.nelmt / 32, EL, HexCoaPipe<T, NQ, EL, NW>::BOXR)) \
            printf("# tensor map encode failed\n");                                                          \
        else                                                                                                 \
            c.run_args("coa-pipe tma EL=" #EL " NW=" #NW " MINB=" #MB, bwdtrans_hex_coapipe_tma_kernel<T, NQ, EL, NW, MB>, 0u, \
                       HexCoaPipe<T, NQ, EL, NW>::THREADS, HexCoaPipe<T, NQ, EL, NW>::SMEM_TMA, tm, c.out, c.nelmt / EL); \
    }
#define CM(NQ, WARPS)                                                                                        \
    c.run_args("coa-mma WARPS=" #WARPS, bwdtrans_quad_coamma_kernel<NQ, WARPS>, 0u, WARPS * 32,                \
               QuadCoaMma<NQ, WARPS>::SMEM, (const double *)c.b[0], (const double *)c.b[1], (const double *)c.in, c.out, c.nelmt / 8);
#define CM32(NQ, WARPS, MB)                                                                                  \
    c.run_args("coa-mma32 WARPS=" #WARPS " MINB=" #MB, bwdtrans_quad_coamma32_kernel<NQ, WARPS, MB>, 0u, WARPS * 32, \
               QuadCoaMma32<NQ, WARPS>::SMEM, (const float *)c.b[0], (const float *)c.b[1], (const float *)c.in, c.out, c.nelmt / 16);

int main(int argc, char **argv)
{
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    printf("op,nq,dtype,_,threads,smem,ctas_per_sm,ms_best,ms_avg,gb_s,mismatches\n");
    if (which == 1)
    {
        {
            Case<double> c;
            c.setup(3, 10);
            HQ(double, 10, 8, 2, 1)
            CP(double, 10, 8, 34, 2) CP(double, 10, 8, 50, 2) CP(double, 10, 8, 52, 2) CP(double, 10, 8, 46, 2) CP(double, 10, 8, 41, 2) CP(double, 10, 8, 27, 2)
            CP(double, 10, 16, 50, 1) CP(double, 10, 16, 25, 1)
            c.teardown();
            c.setup(3, 8);
            HP(double, 8, 16, 1)
            CP(double, 8, 8, 32, 2) CP(double, 8, 8, 32, 4) CP(double, 8, 8, 22, 4) CP(double, 8, 8, 16, 4) CP(double, 8, 16, 32, 2)
            c.teardown();
            c.setup(2, 32);
            QL(double, 32, 16)
            CM(32, 4) CM(32, 2) CM(32, 8)
            c.teardown();
        }
        {
            Case<float> c;
            c.setup(3, 10);
            HP(float, 10, 16, 3)
            CP(float, 10, 16, 25, 2) CP(float, 10, 16, 50, 1) CP(float, 10, 16, 34, 2) CP(float, 10, 32, 25, 1) CP(float, 10, 32, 17, 1)
            c.teardown();
        }
        return 0;
    }
    if (which == 9)
    {
        Case<double> c;
        c.setup(3, 10);
        CPT(double, 10, 8, 52, 2) CPT(double, 10, 8, 50, 2) CPT(double, 10, 8, 54, 2) CPT(double, 10, 8, 56, 2) CPT(double, 10, 8, 34, 2) CPT(double, 10, 8, 27, 2) CPT(double, 10, 8, 60, 2)
        c.teardown();
        c.setup(3, 8);
        CPT(double, 8, 8, 32, 2) CPT(double, 8, 8, 32, 3) CPT(double, 8, 8, 28, 3) CPT(double, 8, 8, 22, 3) CPT(double, 8, 8, 34, 3) CPT(double, 8, 8, 36, 2)
        c.teardown();
        return 0;
    }
    if (which == 3) // FP32 quad nq = 32 on the TF32 tensor-core path (not bit-identical: prints the error)
    {
        Case<float> c;
        c.setup(2, 32);
        QL(float, 32, 16) CM32(32, 4, 2) CM32(32, 4, 3) CM32(32, 3, 3) CM32(32, 2, 3) CM32(32, 6, 2) CM32(32, 5, 3)
        c.teardown();
        return 0;
    }
    if (which == 2) // ncu targets
    {
        Case<double> c;
        c.setup(3, 10);
        CP(double, 10, 8, 50, 2)
        c.teardown();
        c.setup(2, 32);
        CM(32, 4)
        c.teardown();
        Case<float> cf;
        cf.setup(2, 32);
        {
            auto &c = cf;
            CM32(32, 4, 2)
        }
        cf.teardown();
        return 0;
    }
    {
        Case<float> c;
        c.setup(3, 10); HP(float, 10, 16, 3) HP(float, 10, 8, 1) HP(float, 10, 8, 4) HP(float, 10, 8, 5) HP(float, 10, 8, 6) c.teardown();
        c.setup(3, 8); HP(float, 8, 16, 5) HP(float, 8, 8, 1) HP(float, 8, 8, 8) c.teardown();
        c.setup(3, 9); HP(float, 9, 32, 1) HP(float, 9, 8, 5) HP(float, 9, 16, 4) c.teardown();
    }
    {
        Case<double> c;
        c.setup(3, 8); HP(double, 8, 16, 1) HP(double, 8, 8, 1) HP(double, 8, 8, 4) c.teardown();
    }
    return 0;
}

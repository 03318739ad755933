// benchmark03 -- dense matrix-vector product y = A x on B200 through libb200fe.
//
// No CLI arguments, square sizes 128 .. 16384 doubling, reference log format
// (benchmark03/benchmark03.cc:320-336, main :339-350).  A[i][j] = sin(i*N+j+1)
// row-major, x[j] = j (benchmark03.cc:250-258).  Column map:
//   1 "Kokkos"       host CPU, OpenMP row dots
//   2 "cuBLAS (rm)"  gemv 'T' on the row-major matrix   (benchmark03.cc:181-185)
//   3 "cuBLAS (cm)"  gemv 'N' on a column-major copy     (benchmark03.cc:225-229)
//   4 "Cuda"         b200fe_compute_matvec(vl=0)
//   5 "Cuda (vl)"    b200fe_compute_matvec(vl=1)
// Env: B200FE_SIZES  B200FE_DTYPE  B200FE_REPS  B200FE_SKIP_CPU
#include <cublas_v2.h>

#include "../utils/bench_common.h"

using namespace bench;

namespace
{

template <typename T> struct Api;
template <> struct Api<double>
{
    static constexpr auto matvec = b200fe_compute_matvec_f64;
    static constexpr auto sumsq  = b200fe_sumsq_f64;
    static cublasStatus_t gemv(cublasHandle_t h, cublasOperation_t op, int m, int n, const double *A, const double *x,
                               double *y)
    {
        const double one = 1.0, zero = 0.0;
        return cublasDgemv(h, op, m, n, &one, A, m, x, 1, &zero, y, 1);
    }
};
template <> struct Api<float>
{
    static constexpr auto matvec = b200fe_compute_matvec_f32;
    static constexpr auto sumsq  = b200fe_sumsq_f32;
    static cublasStatus_t gemv(cublasHandle_t h, cublasOperation_t op, int m, int n, const float *A, const float *x,
                               float *y)
    {
        const float one = 1.0f, zero = 0.0f;
        return cublasSgemv(h, op, m, n, &one, A, m, x, 1, &zero, y, 1);
    }
};

template <typename T> void run_test(const unsigned size)
{
    using A             = Api<T>;
    const unsigned M = size, N = size;
    const unsigned reps = (unsigned)env_long("B200FE_REPS", 40);
    double secs[5], result[5];
    std::fill(secs, secs + 5, std::numeric_limits<double>::infinity());
    std::fill(result, result + 5, 0.0);

    std::vector<T> h_A((size_t)M * N), h_At((size_t)M * N), h_x(N), h_y(M);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)M; ++i)
        for (unsigned j = 0; j < N; ++j)
        {
            const T v                 = std::sin((T)((unsigned)i * N + j + 1u));
            h_A[(size_t)i * N + j]    = v;
            h_At[(size_t)j * M + i]   = v; // column-major copy
        }
    for (unsigned j = 0; j < N; ++j)
        h_x[j] = (T)j;

    if (!env_long("B200FE_SKIP_CPU", 0))
    {
        secs[0] = time_min_host(std::min(reps, 5u), [&] {
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < (long long)M; ++i)
            {
                T s = 0;
                for (unsigned j = 0; j < N; ++j)
                    s += h_A[(size_t)i * N + j] * h_x[j];
                h_y[i] = s;
            }
        });
        result[0] = host_sumsq(h_y);
    }

    DeviceArray<T> d_A(h_A.size()), d_x(N), d_y(M);
    DeviceArray<double> res(1);
    DeviceArray<unsigned char> scratch(b200fe_sumsq_scratch_bytes());
    d_x.upload(h_x);
    cublasHandle_t handle;
    if (cublasCreate(&handle) != CUBLAS_STATUS_SUCCESS)
        die("cublasCreate", -1, __FILE__, __LINE__);
    for (int col = 1; col < 5; ++col)
    {
        d_A.upload(col == 2 ? h_At : h_A);
        d_y.zero();
        secs[col] = time_min(reps, [&] {
            if (col == 1 || col == 2)
            {
                // row-major M x N seen by cuBLAS as column-major N x M: y = A x is op T (needs M == N like the reference)
                if (A::gemv(handle, col == 1 ? CUBLAS_OP_T : CUBLAS_OP_N, (int)M, (int)N, d_A.get(), d_x.get(),
                            d_y.get()) != CUBLAS_STATUS_SUCCESS)
                    die("cublas gemv", -1, __FILE__, __LINE__);
            }
            else
                FE_OK(A::matvec(N, M, d_A.get(), d_x.get(), d_y.get(), col == 4, nullptr));
        });
        FE_OK(A::sumsq(d_y.get(), M, res.get(), scratch.get(), nullptr));
        CUDA_OK(cudaMemcpy(&result[col], res.get(), sizeof(double), cudaMemcpyDeviceToHost));
    }
    cublasDestroy(handle);

    std::cout << std::setprecision(10);
    std::cout << "Size " << size << " Case:     Kokkos      cuBLAS (rm)   cuBLAS (cm)     Cuda          Cuda (vl)"
              << std::endl;
    std::cout << "Size " << size << " norm:";
    for (int c = 0; c < 5; ++c)
        std::cout << (c ? "     " : " ") << std::sqrt(result[c]);
    std::cout << std::endl;
    const double gb = sizeof(T) * 1.0e-9 * M * N; // the matrix only (benchmark03.cc:332)
    std::cout << "Size " << size << " GB/s:";
    for (int c = 0; c < 5; ++c)
        std::cout << (c ? "     " : " ") << gb / secs[c];
    std::cout << std::endl;
    std::cout << "info " << size << " HBM% of " << hbm_peak_gbs() << ":";
    for (int c = 1; c < 5; ++c)
        std::cout << " " << std::setprecision(4) << 100.0 * gb / secs[c] / hbm_peak_gbs();
    std::cout << (size < 8192u ? " (L2-resident / launch-bound size)" : "") << " | host threads " << host_threads()
              << std::endl;
}

} // namespace

int main(int, char **)
{
    std::cout << "--------------------------------" << std::endl;
    std::cout << "Benchmark03 : Matrix-Vector Mult" << std::endl;
    std::cout << "--------------------------------" << std::endl;
    FE_OK(b200fe_check_device());
    std::vector<unsigned> sizes = env_list("B200FE_SIZES");
    if (sizes.empty())
        for (unsigned size = 2 << 6; size < 2 << 14; size *= 2)
            sizes.push_back(size);
    const std::string dtype = env_str("B200FE_DTYPE", "double");
    for (unsigned size : sizes)
    {
        if (dtype != "float")
            run_test<double>(size);
        if (dtype == "float" || dtype == "both")
            run_test<float>(size);
    }
    return 0;
}

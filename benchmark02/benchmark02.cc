// benchmark02 -- vector addition x += y on B200 through libb200fe.
//
// No CLI arguments, sizes 1024 .. 2^29 doubling, reference log format
// (benchmark02/benchmark02.cc:242-259, main :262-273).  As in the reference
// the timed kernel is applied `reps` = 40 times IN PLACE, so the checksum is
// ||x0 + 40 y|| (benchmark02.cc:152-164).  Column map:
//   1 "Kokkos"          host CPU, OpenMP
//   2 "Thrust"          cuBLAS axpy (vendor library comparator)
//   3 "Cuda"            b200fe_add_vector(vl=0)
//   4 "Cuda (vl)"       b200fe_add_vector(vl=1)
//   5 "Cuda (functor)"  b200fe_vector_kernel_add
// Env: B200FE_SIZES  B200FE_DTYPE  B200FE_REPS (changes the checksum!)  B200FE_SKIP_CPU
#include <cublas_v2.h>

#include "../utils/bench_common.h"

using namespace bench;

namespace
{

template <typename T> struct Api;
template <> struct Api<double>
{
    static constexpr auto set_data  = b200fe_set_data_hostgen_f64; // benchmark02 initialises on the host
    static constexpr auto set_data2 = b200fe_set_data2_f64;
    static constexpr auto add       = b200fe_add_vector_f64;
    static constexpr auto functor   = b200fe_vector_kernel_add_f64;
    static constexpr auto sumsq     = b200fe_sumsq_f64;
    static cublasStatus_t axpy(cublasHandle_t h, int n, const double *y, double *x)
    {
        const double one = 1.0;
        return cublasDaxpy(h, n, &one, y, 1, x, 1);
    }
};
template <> struct Api<float>
{
    static constexpr auto set_data  = b200fe_set_data_hostgen_f32;
    static constexpr auto set_data2 = b200fe_set_data2_f32;
    static constexpr auto add       = b200fe_add_vector_f32;
    static constexpr auto functor   = b200fe_vector_kernel_add_f32;
    static constexpr auto sumsq     = b200fe_sumsq_f32;
    static cublasStatus_t axpy(cublasHandle_t h, int n, const float *y, float *x)
    {
        const float one = 1.0f;
        return cublasSaxpy(h, n, &one, y, 1, x, 1);
    }
};

template <typename T> void run_test(const unsigned size)
{
    using A             = Api<T>;
    const unsigned reps = (unsigned)env_long("B200FE_REPS", 40);
    double secs[5], result[5];
    std::fill(secs, secs + 5, std::numeric_limits<double>::infinity());
    std::fill(result, result + 5, 0.0);

    if (!env_long("B200FE_SKIP_CPU", 0))
    {
        std::vector<T> x(size), y(size);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)size; ++i)
        {
            x[i] = (T)((unsigned)i % 13 + (0.2 + 0.00001 * ((unsigned)i % 100191)));
            y[i] = (T)((unsigned)i % 8 + (0.4 + 0.00003 * ((unsigned)i % 100721)));
        }
        secs[0] = time_min_host(reps, [&] {
#pragma omp parallel for schedule(static)
            for (long long i = 0; i < (long long)size; ++i)
                x[i] += y[i];
        });
        result[0] = host_sumsq(x);
    }

    DeviceArray<T> x(size), y(size);
    DeviceArray<double> res(1);
    DeviceArray<unsigned char> scratch(b200fe_sumsq_scratch_bytes());
    cublasHandle_t handle;
    if (cublasCreate(&handle) != CUBLAS_STATUS_SUCCESS)
        die("cublasCreate", -1, __FILE__, __LINE__);
    FE_OK(A::set_data2(y.get(), size, nullptr));
    for (int col = 1; col < 5; ++col)
    {
        FE_OK(A::set_data(x.get(), size, nullptr)); // x is modified in place: re-initialise per variant
        secs[col] = time_min(reps, [&] {
            if (col == 1)
            {
                if (A::axpy(handle, (int)size, y.get(), x.get()) != CUBLAS_STATUS_SUCCESS)
                    die("cublas axpy", -1, __FILE__, __LINE__);
            }
            else if (col == 4)
                FE_OK(A::functor(0u, size, x.get(), y.get(), nullptr));
            else
                FE_OK(A::add(x.get(), y.get(), size, col == 3, nullptr));
        });
        FE_OK(A::sumsq(x.get(), size, res.get(), scratch.get(), nullptr));
        CUDA_OK(cudaMemcpy(&result[col], res.get(), sizeof(double), cudaMemcpyDeviceToHost));
    }
    cublasDestroy(handle);

    std::cout << std::setprecision(10);
    std::cout << "Size " << size << " Case:     Kokkos      Thrust      Cuda        Cuda (vl)        Cuda (functor)"
              << std::endl;
    std::cout << "Size " << size << " norm: " << std::sqrt(result[0]) << " " << std::sqrt(result[1]) << " "
              << " " << std::sqrt(result[2]) << " " << std::sqrt(result[3]) << " " << std::sqrt(result[4])
              << std::endl;
    const double gb = sizeof(T) * 3e-9 * size; // two reads + one write per element (benchmark02.cc:255)
    std::cout << "Size " << size << " GB/s:";
    for (int c = 0; c < 5; ++c)
        std::cout << " " << gb / secs[c];
    std::cout << std::endl;
    std::cout << "info " << size << " HBM% of " << hbm_peak_gbs() << ":";
    for (int c = 1; c < 5; ++c)
        std::cout << " " << std::setprecision(4) << 100.0 * gb / secs[c] / hbm_peak_gbs();
    std::cout << (2.0 * size * sizeof(T) < (double)((size_t)256 << 20) ? " (L2-resident / launch-bound size)" : "")
              << " | host threads " << host_threads() << std::endl;
}

} // namespace

int main(int, char **)
{
    std::cout << "--------------------------------" << std::endl;
    std::cout << "Benchmark02 : Vector Addition   " << std::endl;
    std::cout << "--------------------------------" << std::endl;
    FE_OK(b200fe_check_device());
    std::vector<unsigned> sizes = env_list("B200FE_SIZES");
    if (sizes.empty())
        for (unsigned size = 1024; size < 1000000000u; size *= 2)
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

// benchmark01 -- L2-norm reduction on B200 through libb200fe.
//
// No CLI arguments, sizes 1024 .. 2^29 doubling, three lines per size in the
// reference's format (benchmark01/benchmark01.cc:317-335, main :337-348), so
// postprocess.py works unchanged.  Column map (5 columns, reference labels):
//   1 "Kokkos"          host CPU, OpenMP reduction
//   2 "Thrust"          the library's deterministic checksum kernels (b200fe_sumsq_*),
//                       the replacement of thrust::transform_reduce
//   3 "CUDA"            b200fe_l2norm_vl(vl=0) + b200fe_reduce_vl
//   4 "CUDA (vl)"       b200fe_l2norm_vl(vl=1) + b200fe_reduce_vl
//   5 "CUDA (functor)"  b200fe_reduceSumKernel_sumsq + b200fe_reduce_vl
// The timed region of columns 3-5 is the reference's (benchmark01.cc:245-252):
// two memsets, two launches and the device-to-host copy of the result.
// Env: B200FE_SIZES=a,b,..  B200FE_DTYPE=double|float|both  B200FE_REPS=40  B200FE_SKIP_CPU=1
#include "../utils/bench_common.h"

using namespace bench;

namespace
{

template <typename T> struct Api;
template <> struct Api<double>
{
    static constexpr auto set_data = b200fe_set_data_f64;
    static constexpr auto l2norm   = b200fe_l2norm_vl_f64;
    static constexpr auto reduce   = b200fe_reduce_vl_f64;
    static constexpr auto functor  = b200fe_reduceSumKernel_sumsq_f64;
    static constexpr auto sumsq    = b200fe_sumsq_f64;
};
template <> struct Api<float>
{
    static constexpr auto set_data = b200fe_set_data_f32;
    static constexpr auto l2norm   = b200fe_l2norm_vl_f32;
    static constexpr auto reduce   = b200fe_reduce_vl_f32;
    static constexpr auto functor  = b200fe_reduceSumKernel_sumsq_f32;
    static constexpr auto sumsq    = b200fe_sumsq_f32;
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
        std::vector<T> h(size);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)size; ++i)
            h[i] = (T)((unsigned)i % 13 + (0.2 + 0.00001 * ((unsigned)i % 100191)));
        T acc   = 0;
        secs[0] = time_min_host(std::min(reps, 10u), [&] {
            T s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
            for (long long i = 0; i < (long long)size; ++i)
                s += h[i] * h[i];
            acc = s;
        });
        result[0] = (double)acc;
    }

    DeviceArray<T> data(size);
    FE_OK(A::set_data(data.get(), size, nullptr));

    {
        DeviceArray<double> res(1);
        DeviceArray<unsigned char> scratch(b200fe_sumsq_scratch_bytes());
        double h = 0.0;
        secs[1]  = time_min(reps, [&] {
            FE_OK(A::sumsq(data.get(), size, res.get(), scratch.get(), nullptr));
            CUDA_OK(cudaMemcpy(&h, res.get(), sizeof(double), cudaMemcpyDeviceToHost));
        });
        result[1] = h;
    }

    const unsigned threads = 256u;
    const unsigned blocks  = std::min((size + threads - 1u) / threads, 1024u); // benchmark01.cc:236-238
    DeviceArray<T> sums(blocks), res(1);
    for (int variant = 0; variant < 3; ++variant)
    {
        T h = 0;
        secs[2 + variant] = time_min(reps, [&] {
            CUDA_OK(cudaMemset(sums.get(), 0, blocks * sizeof(T)));
            CUDA_OK(cudaMemset(res.get(), 0, sizeof(T)));
            if (variant == 2)
                FE_OK(A::functor(0u, size, sums.get(), data.get(), blocks, nullptr));
            else
                FE_OK(A::l2norm(sums.get(), data.get(), size, blocks, variant, nullptr));
            FE_OK(A::reduce(res.get(), sums.get(), blocks, variant != 0, nullptr));
            CUDA_OK(cudaMemcpy(&h, res.get(), sizeof(T), cudaMemcpyDeviceToHost));
        });
        result[2 + variant] = (double)h;
    }

    std::cout << std::setprecision(10);
    std::cout << "Size " << size << " Case:     Kokkos      Thrust      CUDA        CUDA (vl)        CUDA (functor)"
              << std::endl;
    std::cout << "Size " << size << " norm: " << std::sqrt(result[0]) << " " << std::sqrt(result[1]) << " "
              << " " << std::sqrt(result[2]) << " " << std::sqrt(result[3]) << " " << std::sqrt(result[4])
              << std::endl;
    const double gb = sizeof(T) * 1e-9 * size; // one read per element (benchmark01.cc:330)
    std::cout << "Size " << size << " GB/s:";
    for (int c = 0; c < 5; ++c)
        std::cout << " " << gb / secs[c];
    std::cout << std::endl;
    std::cout << "info " << size << " HBM% of " << hbm_peak_gbs() << ":";
    for (int c = 1; c < 5; ++c)
        std::cout << " " << std::setprecision(4) << 100.0 * gb / secs[c] / hbm_peak_gbs();
    std::cout << (size * sizeof(T) < (size_t)256 << 20 ? " (L2-resident / launch-bound size)" : "") << " | host threads "
              << host_threads() << std::endl;
}

} // namespace

int main(int, char **)
{
    std::cout << "--------------------------------" << std::endl;
    std::cout << "Benchmark01 : L2 norm reduction " << std::endl;
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

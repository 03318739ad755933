// benchmark05 -- BwdTrans (3D, hexahedral elements) on B200 through libb200fe.
//
// Same CLI, sweep and log format as the reference driver
// (benchmark05/benchmark05.cc:1423-1442 main, :1387-1420 output):
//
//     benchmark05 [nq0=8] [nq1=8] [nq2=8] [threads=128] [elblocks=1]
//
// Column map and environment knobs as in benchmark04 (see there); column 5 is
// one GEMM + two strided-batched GEMMs on cuBLAS.  Column 7 "Cuda (Coales)"
// carries the CORRECT checksum here: the reference kernel's output offset
// misses a factor nq2 (benchmark05.cc:193), the library implements the
// intended layout (benchmark05.cc:810-812).
#include "../utils/bench_common.h"
#include "../utils/cpu_reference.h"
#include "../utils/cublas_compare.h"
#include "../utils/multi_gpu.h"

using namespace bench;

namespace
{

template <typename T> struct Api;
template <> struct Api<double>
{
    static constexpr auto uncoa = b200fe_BwdTransHexKernel_f64;
    static constexpr auto coa   = b200fe_BwdTransHexKernel_Coa_f64;
    static constexpr auto qp    = b200fe_BwdTransHexKernel_QP_f64;
    static constexpr auto qpsh  = b200fe_BwdTransHexKernel_QP_Shared_f64;
    static constexpr auto q1d   = b200fe_BwdTransHexKernel_QP_1D_f64;
    static constexpr auto q1dsh = b200fe_BwdTransHexKernel_QP_1D_Shared_f64;
    static constexpr const char *name = "double";
};
template <> struct Api<float>
{
    static constexpr auto uncoa = b200fe_BwdTransHexKernel_f32;
    static constexpr auto coa   = b200fe_BwdTransHexKernel_Coa_f32;
    static constexpr auto qp    = b200fe_BwdTransHexKernel_QP_f32;
    static constexpr auto qpsh  = b200fe_BwdTransHexKernel_QP_Shared_f32;
    static constexpr auto q1d   = b200fe_BwdTransHexKernel_QP_1D_f32;
    static constexpr auto q1dsh = b200fe_BwdTransHexKernel_QP_1D_Shared_f32;
    static constexpr const char *name = "float";
};

template <typename T>
void run_test(const unsigned nelmt, const unsigned nq0, const unsigned nq1, const unsigned nq2,
              const unsigned /*threads*/, const unsigned /*elblocks*/)
{
    const unsigned nm0 = nq0 - 1u, nm1 = nq1 - 1u, nm2 = nq2 - 1u;
    const size_t nmTot = (size_t)nm0 * nm1 * nm2, nqTot = (size_t)nq0 * nq1 * nq2;
    const unsigned reps     = (unsigned)env_long("B200FE_REPS", 40);
    const unsigned cpu_reps = (unsigned)env_long("B200FE_CPU_REPS", nelmt > 16384u ? 2 : 5);
    const bool skip_cpu     = env_long("B200FE_SKIP_CPU", 0) != 0;
    const bool skip_blas    = env_long("B200FE_SKIP_CUBLAS", 0) != 0;

    constexpr int kCols = 11;
    double secs[kCols], sumsq[kCols];
    std::fill(secs, secs + kCols, std::numeric_limits<double>::infinity());
    std::fill(sumsq, sumsq + kCols, 0.0);

    // in[e][k] = sin((T)(k+1)), B[k] = cos((T)k)  (benchmark05.cc:1195-1236)
    std::vector<T> h_in(nelmt * nmTot), h_in_coa(nelmt * nmTot), h_b0((size_t)nm0 * nq0), h_b1((size_t)nm1 * nq1),
        h_b2((size_t)nm2 * nq2);
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < (long long)nelmt; ++e)
        for (size_t k = 0; k < nmTot; ++k)
        {
            const T v = std::sin((T)(k + 1u));
            h_in[(size_t)e * nmTot + k]                                          = v;
            h_in_coa[cpuref::at(cpuref::Layout::Interleaved, (size_t)e, k, nmTot)] = v;
        }
    for (size_t k = 0; k < h_b0.size(); ++k)
        h_b0[k] = std::cos((T)k);
    for (size_t k = 0; k < h_b1.size(); ++k)
        h_b1[k] = std::cos((T)k);
    for (size_t k = 0; k < h_b2.size(); ++k)
        h_b2[k] = std::cos((T)k);

    if (!skip_cpu)
    {
        std::vector<T> h_out(nelmt * nqTot);
        secs[0] = time_min_host(cpu_reps, [&] {
            cpuref::bwdtrans_hex_nest(cpuref::Layout::ElementMajor, nm0, nm1, nm2, nq0, nq1, nq2, nelmt, h_b0.data(),
                                      h_b1.data(), h_b2.data(), h_in.data(), h_out.data());
        });
        sumsq[0] = host_sumsq(h_out);
        secs[1] = time_min_host(cpu_reps, [&] {
            cpuref::bwdtrans_hex_nest(cpuref::Layout::Interleaved, nm0, nm1, nm2, nq0, nq1, nq2, nelmt, h_b0.data(),
                                      h_b1.data(), h_b2.data(), h_in_coa.data(), h_out.data());
        });
        sumsq[1] = host_sumsq(h_out);
        secs[2] = time_min_host(cpu_reps, [&] {
            cpuref::bwdtrans_hex_threepass(nm0, nm1, nm2, nq0, nq1, nq2, nelmt, h_b0.data(), h_b1.data(), h_b2.data(),
                                           h_in.data(), h_out.data());
        });
        sumsq[2] = host_sumsq(h_out);
#ifdef _OPENMP
        const int nthreads = omp_get_max_threads();
        omp_set_num_threads(1);
#endif
        const size_t n1 = std::min<size_t>(nelmt, 2048);
        secs[3] = time_min_host(1, [&] {
            cpuref::bwdtrans_hex_threepass(nm0, nm1, nm2, nq0, nq1, nq2, n1, h_b0.data(), h_b1.data(), h_b2.data(),
                                           h_in.data(), h_out.data());
        }) * ((double)nelmt / (double)n1);
        sumsq[3] = sumsq[2];
#ifdef _OPENMP
        omp_set_num_threads(nthreads);
#endif
    }

    // caller-owned device buffers, scratch included (benchmark05.cc:1237-1258)
    DeviceArray<T> d_in(nelmt * nmTot), d_in_coa(nelmt * nmTot), d_out(nelmt * nqTot), d_b0(h_b0.size()),
        d_b1(h_b1.size()), d_b2(h_b2.size()), d_wsp1((size_t)nelmt * nq0 * nm1 * nm2),
        d_wsp2((size_t)nelmt * nq0 * nq1 * nm2);
    d_in.upload(h_in);
    d_in_coa.upload(h_in_coa);
    d_b0.upload(h_b0);
    d_b1.upload(h_b1);
    d_b2.upload(h_b2);
    Checksum<T> checksum;

    if (!skip_blas)
    {
        cublasHandle_t handle;
        CUBLAS_OK(cublasCreate(&handle));
        d_out.zero();
        secs[4] = time_min(reps, [&] {
            blascmp::bwdtrans_hex<T>(handle, (int)nm0, (int)nm1, (int)nm2, (int)nq0, (int)nq1, (int)nq2, (int)nelmt,
                                     d_b0.get(), d_b1.get(), d_b2.get(), d_in.get(), d_wsp1.get(), d_wsp2.get(),
                                     d_out.get());
        });
        sumsq[4] = checksum(d_out.get(), d_out.size());
        CUBLAS_OK(cublasDestroy(handle));
    }
    // B200FE_COL5=gemm: the same GEMM factorisation (intermediates in global memory) on the library's own kernels
    // instead of cuBLAS -- a cuBLAS-free column 5
    if (env_str("B200FE_COL5", "cublas") == "gemm")
    {
        d_out.zero();
        secs[4] = time_min(reps, [&] {
            if constexpr (std::is_same<T, double>::value)
                FE_OK(b200fe_gemm_bwdtrans_hex_f64(nm0, nm1, nm2, nq0, nq1, nq2, nelmt, d_b0.get(), d_b1.get(), d_b2.get(),
                                                   d_in.get(), d_wsp1.get(), d_wsp2.get(), d_out.get(), nullptr));
            else
                FE_OK(b200fe_gemm_bwdtrans_hex_f32(nm0, nm1, nm2, nq0, nq1, nq2, nelmt, d_b0.get(), d_b1.get(), d_b2.get(),
                                                   d_in.get(), d_wsp1.get(), d_wsp2.get(), d_out.get(), nullptr));
        });
        sumsq[4] = checksum(d_out.get(), d_out.size());
    }

    using A = Api<T>;
    const OperatorPlan plan(3, sizeof(T) == 4, nq0, nq1, nq2, d_b0.get(), d_b1.get(), d_b2.get());
    auto with_wsp = [&](auto fn, const T *in, bool interleaved = false) {
        if (plan.active())
            return plan.bwdtrans(interleaved, nelmt, in, d_out.get());
        FE_OK(fn(nm0, nm1, nm2, (unsigned)nmTot, nq0, nq1, nq2, nelmt, d_b0.get(), d_b1.get(), d_b2.get(), in,
                 d_wsp1.get(), d_wsp2.get(), d_out.get(), nullptr));
    };
    auto no_wsp = [&](auto fn) {
        if (plan.active())
            return plan.bwdtrans(false, nelmt, d_in.get(), d_out.get());
        FE_OK(fn(nm0, nm1, nm2, (unsigned)nmTot, nq0, nq1, nq2, nelmt, d_b0.get(), d_b1.get(), d_b2.get(), d_in.get(),
                 d_out.get(), nullptr));
    };
    auto column = [&](int col, auto &&launch) {
        d_out.zero();
        secs[col]  = time_min(reps, launch);
        sumsq[col] = checksum(d_out.get(), d_out.size());
    };
    column(5, [&] { with_wsp(A::uncoa, d_in.get()); });
    if (nelmt % 32u == 0)
        column(6, [&] { with_wsp(A::coa, d_in_coa.get(), true); });
    column(7, [&] { with_wsp(A::qp, d_in.get()); });
    column(8, [&] { no_wsp(A::qpsh); });
    column(9, [&] { with_wsp(A::q1d, d_in.get()); });
    column(10, [&] { no_wsp(A::q1dsh); });

    std::cout << std::setprecision(10);
    std::cout << "nelmt " << nelmt
              << " Case: Kokkos (Uncoales) Kokkos (Coales) Kokkos (QP)   Kokkos (QP/Shared) cuBLAS       Cuda "
                 "(Uncoales) Cuda (Coales)    Cuda (QP)      Cuda (QP/Shared)  Cuda (QP-1D)   Cuda (QP-1D/Shared)"
              << std::endl;
    std::cout << "nelmt " << nelmt << " norm:";
    for (int c = 0; c < kCols; ++c)
        std::cout << (c ? "     " : " ") << std::sqrt(sumsq[c]);
    std::cout << std::endl;
    const double dof = 1.0e-9 * (double)nelmt * (double)nmTot;
    std::cout << "nelmt " << nelmt << " DOF/s:";
    for (int c = 0; c < kCols; ++c)
        std::cout << (c ? "     " : " ") << dof / secs[c];
    std::cout << std::endl;

    const double gb = 1.0e-9 * (double)nelmt * (double)sizeof(T) * (double)(nmTot + nqTot);
    std::cout << "info " << nelmt << " " << A::name << " HBM% of " << hbm_peak_gbs() << " GB/s, columns 6-11:";
    for (int c = 5; c < kCols; ++c)
        std::cout << " " << std::setprecision(4) << 100.0 * gb / secs[c] / hbm_peak_gbs();
    std::cout << " | host threads " << host_threads() << (plan.active() ? " | via b200fe_plan" : " | per-call entry points")
              << std::endl
              << std::flush;
}

// B200FE_NGPUS > 1: the element range is sharded over the GPUs (one host thread each); the six Cuda columns
// report whole-job throughput (slowest GPU), the norm is all-reduced with NCCL.  Host / cuBLAS columns: 0.
template <typename T>
void run_test_multi(MultiGpu &mg, const unsigned nelmt, const unsigned nq0, const unsigned nq1, const unsigned nq2)
{
    const unsigned nm0 = nq0 - 1u, nm1 = nq1 - 1u, nm2 = nq2 - 1u;
    const size_t nmTot = (size_t)nm0 * nm1 * nm2, nqTot = (size_t)nq0 * nq1 * nq2;
    const unsigned reps = (unsigned)env_long("B200FE_REPS", 40);
    constexpr int kCols = 11;
    double secs[kCols], sumsq[kCols];
    std::fill(secs, secs + kCols, std::numeric_limits<double>::infinity());
    std::fill(sumsq, sumsq + kCols, 0.0);

    mg.run([&](int rank) {
        size_t e_begin, e_end;
        shard_range(nelmt, rank, mg.size(), 32, &e_begin, &e_end);
        const unsigned n = (unsigned)(e_end - e_begin);
        // every element carries the same modes, so a shard's input does not depend on its position
        std::vector<T> h_in(n * nmTot), h_in_coa(n * nmTot), h_b0((size_t)nm0 * nq0), h_b1((size_t)nm1 * nq1),
            h_b2((size_t)nm2 * nq2);
        for (size_t e = 0; e < n; ++e)
            for (size_t k = 0; k < nmTot; ++k)
            {
                const T v = std::sin((T)(k + 1u));
                h_in[e * nmTot + k] = v;
                if (n % 32u == 0)
                    h_in_coa[cpuref::at(cpuref::Layout::Interleaved, e, k, nmTot)] = v;
            }
        for (size_t k = 0; k < h_b0.size(); ++k)
            h_b0[k] = std::cos((T)k);
        for (size_t k = 0; k < h_b1.size(); ++k)
            h_b1[k] = std::cos((T)k);
        for (size_t k = 0; k < h_b2.size(); ++k)
            h_b2[k] = std::cos((T)k);
        DeviceArray<T> d_in(n * nmTot), d_in_coa(n * nmTot), d_out(n * nqTot), d_b0(h_b0.size()), d_b1(h_b1.size()),
            d_b2(h_b2.size()), d_wsp1((size_t)n * nq0 * nm1 * nm2), d_wsp2((size_t)n * nq0 * nq1 * nm2);
        if (n)
        {
            d_in.upload(h_in);
            d_in_coa.upload(h_in_coa);
        }
        d_b0.upload(h_b0);
        d_b1.upload(h_b1);
        d_b2.upload(h_b2);
        Checksum<T> checksum;

        using A = Api<T>;
        const OperatorPlan plan(3, sizeof(T) == 4, nq0, nq1, nq2, d_b0.get(), d_b1.get(), d_b2.get()); // this rank's device
        auto with_wsp = [&](auto fn, const T *in, bool interleaved = false) {
            if (plan.active())
                return plan.bwdtrans(interleaved, n, in, d_out.get());
            FE_OK(fn(nm0, nm1, nm2, (unsigned)nmTot, nq0, nq1, nq2, n, d_b0.get(), d_b1.get(), d_b2.get(), in,
                     d_wsp1.get(), d_wsp2.get(), d_out.get(), nullptr));
        };
        auto no_wsp = [&](auto fn) {
            if (plan.active())
                return plan.bwdtrans(false, n, d_in.get(), d_out.get());
            FE_OK(fn(nm0, nm1, nm2, (unsigned)nmTot, nq0, nq1, nq2, n, d_b0.get(), d_b1.get(), d_b2.get(), d_in.get(),
                     d_out.get(), nullptr));
        };
        auto column = [&](int col, auto &&launch) {
            if (n)
                d_out.zero();
            const double t = mg.time_min(rank, reps, [&] {
                if (n)
                    launch();
            });
            checksum.launch(d_out.get(), n ? d_out.size() : 0);
            const double total = mg.allreduce_sum(rank, checksum.result.get()); // NCCL: one double
            if (rank == 0)
            {
                secs[col]  = t;
                sumsq[col] = total;
            }
        };
        column(5, [&] { with_wsp(A::uncoa, d_in.get()); });
        bool all32 = true; // the interleaved layout needs every shard to hold whole groups of 32
        for (int r = 0; r < mg.size(); ++r)
        {
            size_t b, e;
            shard_range(nelmt, r, mg.size(), 32, &b, &e);
            all32 = all32 && ((e - b) % 32u == 0);
        }
        if (all32)
            column(6, [&] { with_wsp(A::coa, d_in_coa.get(), true); });
        column(7, [&] { with_wsp(A::qp, d_in.get()); });
        column(8, [&] { no_wsp(A::qpsh); });
        column(9, [&] { with_wsp(A::q1d, d_in.get()); });
        column(10, [&] { no_wsp(A::q1dsh); });
    });

    std::cout << std::setprecision(10);
    std::cout << "nelmt " << nelmt
              << " Case: Kokkos (Uncoales) Kokkos (Coales) Kokkos (QP)   Kokkos (QP/Shared) cuBLAS       Cuda "
                 "(Uncoales) Cuda (Coales)    Cuda (QP)      Cuda (QP/Shared)  Cuda (QP-1D)   Cuda (QP-1D/Shared)"
              << std::endl;
    std::cout << "nelmt " << nelmt << " norm:";
    for (int c = 0; c < kCols; ++c)
        std::cout << (c ? "     " : " ") << std::sqrt(sumsq[c]);
    std::cout << std::endl;
    const double dof = 1.0e-9 * (double)nelmt * (double)nmTot;
    std::cout << "nelmt " << nelmt << " DOF/s:";
    for (int c = 0; c < kCols; ++c)
        std::cout << (c ? "     " : " ") << dof / secs[c];
    std::cout << std::endl;
    const double gb = 1.0e-9 * (double)nelmt * (double)sizeof(T) * (double)(nmTot + nqTot);
    std::cout << "info " << nelmt << " " << Api<T>::name << " gpus " << mg.size() << " aggregate HBM% of "
              << mg.size() << " x " << hbm_peak_gbs() << " GB/s, columns 6-11:";
    for (int c = 5; c < kCols; ++c)
        std::cout << " " << std::setprecision(4) << 100.0 * gb / secs[c] / (hbm_peak_gbs() * mg.size());
    std::cout << std::endl << std::flush;
}

} // namespace

int main(int argc, char **argv)
{
    const unsigned nq0      = (argc > 1) ? (unsigned)atoi(argv[1]) : 8u;
    const unsigned nq1      = (argc > 2) ? (unsigned)atoi(argv[2]) : 8u;
    const unsigned nq2      = (argc > 3) ? (unsigned)atoi(argv[3]) : 8u;
    const unsigned threads  = (argc > 4) ? (unsigned)atoi(argv[4]) : 128u;
    const unsigned elblocks = (argc > 5) ? (unsigned)atoi(argv[5]) : 1u;

    std::cout << "--------------------------------" << std::endl;
    std::cout << "Benchmark05 : BwdTrans (3D)     " << std::endl;
    std::cout << "--------------------------------" << std::endl;
    std::cout << "BwdTrans (NQ = " << nq0 << ", " << nq1 << ", " << nq2 << ")" << std::endl;
    if (nq0 < 2u || nq1 < 2u || nq2 < 2u)
    {
        std::cerr << "benchmark05: nq must be >= 2" << std::endl;
        return 1;
    }
    FE_OK(b200fe_check_device());
    std::cout << "info library " << b200fe_version() << ", host threads " << host_threads() << std::endl;

    std::vector<unsigned> sizes = env_list("B200FE_NELMT");
    if (sizes.empty())
        for (unsigned size = 2 << 6; size < 2 << 20; size <<= 1)
            sizes.push_back(size);
    const std::string dtype = env_str("B200FE_DTYPE", "double");
    const int ngpus         = (int)env_long("B200FE_NGPUS", 1);
    if (ngpus > 1)
    {
        MultiGpu mg(ngpus);
        std::cout << "info sharded over " << ngpus << " GPUs, NCCL scalar all-reduce of the norm only" << std::endl;
        for (unsigned size : sizes)
        {
            if (dtype != "float")
                run_test_multi<double>(mg, size, nq0, nq1, nq2);
            if (dtype == "float" || dtype == "both")
                run_test_multi<float>(mg, size, nq0, nq1, nq2);
        }
        return 0;
    }
    for (unsigned size : sizes)
    {
        if (dtype != "float")
            run_test<double>(size, nq0, nq1, nq2, threads, elblocks);
        if (dtype == "float" || dtype == "both")
            run_test<float>(size, nq0, nq1, nq2, threads, elblocks);
    }
    return 0;
}

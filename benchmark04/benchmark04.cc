// benchmark04 -- BwdTrans (2D, quadrilateral elements) on B200 through libb200fe.
//
// Same CLI, same sweep and same three-lines-per-size log as the reference driver
// (benchmark04/benchmark04.cc:1058-1075 main, :1022-1055 output), so the
// reference's run.sh and postprocess.py work on it unchanged:
//
//     benchmark04 [nq0=8] [nq1=8] [threads=128] [elblocks=1]
//
// `threads` and `elblocks` are accepted for compatibility; launch shapes are
// chosen inside the library.  Column map (11 columns, reference labels kept):
//   1 "Kokkos (Uncoales)"  host CPU, OpenMP, per-column nest, element-major
//   2 "Kokkos (Coales)"    host CPU, OpenMP, per-column nest, interleaved layout
//   3 "Kokkos (QP)"        host CPU, OpenMP, two-pass sum factorisation
//   4 "Kokkos (QP/Shared)" host CPU, single thread, two-pass (per-core rate)
//   5 "cuBLAS"             GEMM + strided-batched GEMM on cuBLAS, as in the reference
//   6..11 "Cuda (...)"     the six entry points of include/b200fe.h
// Extra knobs are environment variables so the positional CLI stays intact:
//   B200FE_NELMT=a,b,..  element counts instead of the 128..1Mi sweep
//   B200FE_DTYPE=double|float|both     B200FE_REPS=40     B200FE_CPU_REPS=2
//   B200FE_SKIP_CPU=1  B200FE_SKIP_CUBLAS=1
//   B200FE_COL5=gemm  column 5 = the GEMM formulation on the library's own kernels (b200fe_gemm_bwdtrans_*) instead of cuBLAS
//   B200FE_PLAN=0  issue every repetition through the per-call entry points instead of a b200fe_plan
// Lines starting with "info" carry roofline figures; postprocess.py ignores them.
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
    static constexpr auto uncoa = b200fe_BwdTransQuadKernel_f64;
    static constexpr auto coa   = b200fe_BwdTransQuadKernel_Coa_f64;
    static constexpr auto qp    = b200fe_BwdTransQuadKernel_QP_f64;
    static constexpr auto qpsh  = b200fe_BwdTransQuadKernel_QP_Shared_f64;
    static constexpr auto q1d   = b200fe_BwdTransQuadKernel_QP_1D_f64;
    static constexpr auto q1dsh = b200fe_BwdTransQuadKernel_QP_1D_Shared_f64;
    static constexpr const char *name = "double";
};
template <> struct Api<float>
{
    static constexpr auto uncoa = b200fe_BwdTransQuadKernel_f32;
    static constexpr auto coa   = b200fe_BwdTransQuadKernel_Coa_f32;
    static constexpr auto qp    = b200fe_BwdTransQuadKernel_QP_f32;
    static constexpr auto qpsh  = b200fe_BwdTransQuadKernel_QP_Shared_f32;
    static constexpr auto q1d   = b200fe_BwdTransQuadKernel_QP_1D_f32;
    static constexpr auto q1dsh = b200fe_BwdTransQuadKernel_QP_1D_Shared_f32;
    static constexpr const char *name = "float";
};

template <typename T>
void run_test(const unsigned nelmt, const unsigned nq0, const unsigned nq1, const unsigned /*threads*/,
              const unsigned /*elblocks*/)
{
    const unsigned nm0 = nq0 - 1u, nm1 = nq1 - 1u;
    const size_t nmTot = (size_t)nm0 * nm1, nqTot = (size_t)nq0 * nq1;
    const unsigned reps     = (unsigned)env_long("B200FE_REPS", 40);
    const unsigned cpu_reps = (unsigned)env_long("B200FE_CPU_REPS", nelmt > 65536u ? 2 : 5);
    const bool skip_cpu     = env_long("B200FE_SKIP_CPU", 0) != 0;
    const bool skip_blas    = env_long("B200FE_SKIP_CUBLAS", 0) != 0;

    constexpr int kCols = 11;
    double secs[kCols], sumsq[kCols];
    std::fill(secs, secs + kCols, std::numeric_limits<double>::infinity());
    std::fill(sumsq, sumsq + kCols, 0.0);

    // synthetic input of the reference: in[e][k] = sin((T)(k+1)), B[k] = cos((T)k)
    // (benchmark04.cc:859-889), in both layouts
    std::vector<T> h_in(nelmt * nmTot), h_in_coa(nelmt * nmTot), h_b0((size_t)nm0 * nq0), h_b1((size_t)nm1 * nq1);
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

    // ---- columns 1-4: the same operator on the host cores ---------------------------------
    if (!skip_cpu)
    {
        std::vector<T> h_out(nelmt * nqTot);
        secs[0] = time_min_host(cpu_reps, [&] {
            cpuref::bwdtrans_quad_nest(cpuref::Layout::ElementMajor, nm0, nm1, nq0, nq1, nelmt, h_b0.data(),
                                       h_b1.data(), h_in.data(), h_out.data());
        });
        sumsq[0] = host_sumsq(h_out);
        secs[1] = time_min_host(cpu_reps, [&] {
            cpuref::bwdtrans_quad_nest(cpuref::Layout::Interleaved, nm0, nm1, nq0, nq1, nelmt, h_b0.data(),
                                       h_b1.data(), h_in_coa.data(), h_out.data());
        });
        sumsq[1] = host_sumsq(h_out);
        secs[2] = time_min_host(cpu_reps, [&] {
            cpuref::bwdtrans_quad_twopass(nm0, nm1, nq0, nq1, nelmt, h_b0.data(), h_b1.data(), h_in.data(),
                                          h_out.data());
        });
        sumsq[2] = host_sumsq(h_out);
#ifdef _OPENMP
        const int nthreads = omp_get_max_threads();
        omp_set_num_threads(1);
#endif
        const size_t n1 = std::min<size_t>(nelmt, 16384); // single core: bounded sample, rate scaled
        secs[3] = time_min_host(1, [&] {
            cpuref::bwdtrans_quad_twopass(nm0, nm1, nq0, nq1, n1, h_b0.data(), h_b1.data(), h_in.data(),
                                          h_out.data());
        }) * ((double)nelmt / (double)n1);
        sumsq[3] = sumsq[2];
#ifdef _OPENMP
        omp_set_num_threads(nthreads);
#endif
    }

    // ---- device buffers (caller-owned, as in the reference: benchmark04.cc:890-905) ---------
    DeviceArray<T> d_in(nelmt * nmTot), d_in_coa(nelmt * nmTot), d_out(nelmt * nqTot), d_b0(h_b0.size()),
        d_b1(h_b1.size()), d_wsp((size_t)nelmt * nq0 * nm1);
    d_in.upload(h_in);
    d_in_coa.upload(h_in_coa);
    d_b0.upload(h_b0);
    d_b1.upload(h_b1);
    Checksum<T> checksum;

    // ---- column 5: cuBLAS ---------------------------------------------------------------------
    if (!skip_blas)
    {
        cublasHandle_t handle;
        CUBLAS_OK(cublasCreate(&handle));
        d_out.zero();
        secs[4] = time_min(reps, [&] {
            blascmp::bwdtrans_quad<T>(handle, (int)nm0, (int)nm1, (int)nq0, (int)nq1, (int)nelmt, d_b0.get(),
                                      d_b1.get(), d_in.get(), d_wsp.get(), d_out.get());
        });
        sumsq[4] = checksum(d_out.get(), d_out.size());
        CUBLAS_OK(cublasDestroy(handle));
    }
    // B200FE_COL5=gemm: the same GEMM factorisation (intermediate in global memory) on the library's own kernels
    // instead of cuBLAS -- a cuBLAS-free column 5
    if (env_str("B200FE_COL5", "cublas") == "gemm")
    {
        d_out.zero();
        secs[4] = time_min(reps, [&] {
            if constexpr (std::is_same<T, double>::value)
                FE_OK(b200fe_gemm_bwdtrans_quad_f64(nm0, nm1, nq0, nq1, nelmt, d_b0.get(), d_b1.get(), d_in.get(),
                                                    d_wsp.get(), d_out.get(), nullptr));
            else
                FE_OK(b200fe_gemm_bwdtrans_quad_f32(nm0, nm1, nq0, nq1, nelmt, d_b0.get(), d_b1.get(), d_in.get(),
                                                    d_wsp.get(), d_out.get(), nullptr));
        });
        sumsq[4] = checksum(d_out.get(), d_out.size());
    }

    // ---- columns 6-11: the six kernel entry points ----------------------------------------------
    using A = Api<T>;
    const OperatorPlan plan(2, sizeof(T) == 4, nq0, nq1, 0u, d_b0.get(), d_b1.get(), nullptr);
    auto with_wsp = [&](auto fn, const T *in, bool interleaved = false) {
        if (plan.active())
            return plan.bwdtrans(interleaved, nelmt, in, d_out.get());
        FE_OK(fn(nm0, nm1, (unsigned)nmTot, nq0, nq1, nelmt, d_b0.get(), d_b1.get(), in, d_wsp.get(), d_out.get(),
                 nullptr));
    };
    auto no_wsp = [&](auto fn) {
        if (plan.active())
            return plan.bwdtrans(false, nelmt, d_in.get(), d_out.get());
        FE_OK(fn(nm0, nm1, (unsigned)nmTot, nq0, nq1, nelmt, d_b0.get(), d_b1.get(), d_in.get(), d_out.get(), nullptr));
    };
    auto column = [&](int col, auto &&launch) {
        d_out.zero(); // a variant can never inherit the previous variant's output
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

    // ---- the reference's three lines (benchmark04.cc:1023-1054) -------------------------------------
    std::cout << std::setprecision(10);
    std::cout << "nelmt " << nelmt
              << " Case: Kokkos (Uncoales) Kokkos (Coales) Kokkos (QP)   Kokkos (QP/Shared) cuBLAS          Cuda "
                 "(Uncoales) Cuda (Coales)    Cuda (QP)      Cuda (QP/Shared)  Cuda (QP-1D)   Cuda (QP-1D/Shared)"
              << std::endl;
    std::cout << "nelmt " << nelmt << " norm:";
    for (int c = 0; c < kCols; ++c)
        std::cout << (c ? "     " : " ") << std::sqrt(sumsq[c]);
    std::cout << std::endl;
    const double dof = 1.0e-9 * (double)nelmt * (double)nmTot; // modes, as in the reference
    std::cout << "nelmt " << nelmt << " DOF/s:";
    for (int c = 0; c < kCols; ++c)
        std::cout << (c ? "     " : " ") << dof / secs[c];
    std::cout << std::endl;

    // ---- roofline side line: algorithmic bytes = sizeof(T) * (modes + points) per element --------
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
template <typename T> void run_test_multi(MultiGpu &mg, const unsigned nelmt, const unsigned nq0, const unsigned nq1)
{
    const unsigned nm0 = nq0 - 1u, nm1 = nq1 - 1u;
    const size_t nmTot = (size_t)nm0 * nm1, nqTot = (size_t)nq0 * nq1;
    const unsigned reps = (unsigned)env_long("B200FE_REPS", 40);
    constexpr int kCols = 11;
    double secs[kCols], sumsq[kCols];
    std::fill(secs, secs + kCols, std::numeric_limits<double>::infinity());
    std::fill(sumsq, sumsq + kCols, 0.0);
    bool all32 = true; // the interleaved layout needs every shard to hold whole groups of 32
    for (int r = 0; r < mg.size(); ++r)
    {
        size_t b, e;
        shard_range(nelmt, r, mg.size(), 32, &b, &e);
        all32 = all32 && ((e - b) % 32u == 0);
    }

    mg.run([&](int rank) {
        size_t e_begin, e_end;
        shard_range(nelmt, rank, mg.size(), 32, &e_begin, &e_end);
        const unsigned n = (unsigned)(e_end - e_begin);
        // every element carries the same modes, so a shard's input does not depend on its position
        std::vector<T> h_in(n * nmTot), h_in_coa(n * nmTot), h_b0((size_t)nm0 * nq0), h_b1((size_t)nm1 * nq1);
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
        DeviceArray<T> d_in(n * nmTot), d_in_coa(n * nmTot), d_out(n * nqTot), d_b0(h_b0.size()), d_b1(h_b1.size()),
            d_wsp((size_t)n * nq0 * nm1);
        if (n)
        {
            d_in.upload(h_in);
            d_in_coa.upload(h_in_coa);
        }
        d_b0.upload(h_b0);
        d_b1.upload(h_b1);
        Checksum<T> checksum;

        using A = Api<T>;
        const OperatorPlan plan(2, sizeof(T) == 4, nq0, nq1, 0u, d_b0.get(), d_b1.get(), nullptr); // on this rank's device
        auto with_wsp = [&](auto fn, const T *in, bool interleaved = false) {
            if (plan.active())
                return plan.bwdtrans(interleaved, n, in, d_out.get());
            FE_OK(fn(nm0, nm1, (unsigned)nmTot, nq0, nq1, n, d_b0.get(), d_b1.get(), in, d_wsp.get(), d_out.get(),
                     nullptr));
        };
        auto no_wsp = [&](auto fn) {
            if (plan.active())
                return plan.bwdtrans(false, n, d_in.get(), d_out.get());
            FE_OK(fn(nm0, nm1, (unsigned)nmTot, nq0, nq1, n, d_b0.get(), d_b1.get(), d_in.get(), d_out.get(), nullptr));
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
        if (all32)
            column(6, [&] { with_wsp(A::coa, d_in_coa.get(), true); });
        column(7, [&] { with_wsp(A::qp, d_in.get()); });
        column(8, [&] { no_wsp(A::qpsh); });
        column(9, [&] { with_wsp(A::q1d, d_in.get()); });
        column(10, [&] { no_wsp(A::q1dsh); });
    });

    std::cout << std::setprecision(10);
    std::cout << "nelmt " << nelmt
              << " Case: Kokkos (Uncoales) Kokkos (Coales) Kokkos (QP)   Kokkos (QP/Shared) cuBLAS          Cuda "
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
    const unsigned threads  = (argc > 3) ? (unsigned)atoi(argv[3]) : 128u;
    const unsigned elblocks = (argc > 4) ? (unsigned)atoi(argv[4]) : 1u;

    std::cout << "--------------------------------" << std::endl;
    std::cout << "Benchmark04 : BwdTrans (2D)     " << std::endl;
    std::cout << "--------------------------------" << std::endl;
    std::cout << "BwdTrans (NQ = " << nq0 << ", " << nq1 << ")" << std::endl;
    if (nq0 < 2u || nq1 < 2u)
    {
        std::cerr << "benchmark04: nq must be >= 2" << std::endl;
        return 1;
    }
    FE_OK(b200fe_check_device());
    std::cout << "info library " << b200fe_version() << ", host threads " << host_threads() << std::endl;

    std::vector<unsigned> sizes = env_list("B200FE_NELMT");
    if (sizes.empty())
        for (unsigned size = 2 << 6; size < 2 << 20; size <<= 1)
            sizes.push_back(size);
    const std::string dtype = env_str("B200FE_DTYPE", "double");
    const int ngpus = (int)env_long("B200FE_NGPUS", 1);
    if (ngpus > 1)
    {
        MultiGpu mg(ngpus);
        std::cout << "info sharded over " << ngpus << " GPUs, NCCL scalar all-reduce of the norm only" << std::endl;
        for (unsigned size : sizes)
        {
            if (dtype != "float")
                run_test_multi<double>(mg, size, nq0, nq1);
            if (dtype == "float" || dtype == "both")
                run_test_multi<float>(mg, size, nq0, nq1);
        }
        return 0;
    }
    for (unsigned size : sizes)
    {
        if (dtype != "float")
            run_test<double>(size, nq0, nq1, threads, elblocks);
        if (dtype == "float" || dtype == "both")
            run_test<float>(size, nq0, nq1, threads, elblocks);
    }
    return 0;
}

#pragma once
// bench_common.h -- host-side helpers shared by the five benchmark drivers
// (plain C++17, no CUDA syntax: the drivers talk to the GPU only through the
// C ABI of libb200fe.so and the CUDA runtime API).
#include <cuda_runtime_api.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/b200fe.h"
#include "timer.h"

namespace bench
{

inline void die(const char *what, int code, const char *file, int line)
{
    std::fprintf(stderr, "%s:%d: %s failed with code %d%s%s\n", file, line, what, code, code > 0 ? ": " : "",
                 code > 0 ? cudaGetErrorString((cudaError_t)code) : "");
    std::exit(2);
}

// every CUDA runtime call and every library call is checked (the reference checks none)
#define CUDA_OK(expr)                                                                                        \
    do                                                                                                       \
    {                                                                                                        \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            bench::die(#expr, (int)e__, __FILE__, __LINE__);                                                 \
    } while (0)
#define FE_OK(expr)                                                                                          \
    do                                                                                                       \
    {                                                                                                        \
        int rc__ = (expr);                                                                                   \
        if (rc__ != 0)                                                                                       \
            bench::die(#expr, rc__, __FILE__, __LINE__);                                                     \
    } while (0)

template <typename T> class DeviceArray
{
public:
    DeviceArray() = default;
    explicit DeviceArray(size_t n) { resize(n); }
    DeviceArray(const DeviceArray &)            = delete;
    DeviceArray &operator=(const DeviceArray &) = delete;
    ~DeviceArray() { release(); }
    void resize(size_t n)
    {
        release();
        m_n = n;
        if (n)
            CUDA_OK(cudaMalloc((void **)&m_p, n * sizeof(T)));
    }
    void release()
    {
        if (m_p)
            cudaFree(m_p);
        m_p = nullptr;
        m_n = 0;
    }
    void upload(const std::vector<T> &h) { CUDA_OK(cudaMemcpy(m_p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); }
    void download(std::vector<T> &h) const
    {
        h.resize(m_n);
        CUDA_OK(cudaMemcpy(h.data(), m_p, m_n * sizeof(T), cudaMemcpyDeviceToHost));
    }
    void zero() { CUDA_OK(cudaMemset(m_p, 0, m_n * sizeof(T))); }
    T *get() const { return m_p; }
    size_t size() const { return m_n; }

private:
    T *m_p     = nullptr;
    size_t m_n = 0;
};

inline std::string env_str(const char *name, const char *dflt)
{
    const char *v = std::getenv(name);
    return v && *v ? std::string(v) : std::string(dflt);
}

inline long env_long(const char *name, long dflt)
{
    const char *v = std::getenv(name);
    return v && *v ? std::atol(v) : dflt;
}

inline double env_double(const char *name, double dflt)
{
    const char *v = std::getenv(name);
    return v && *v ? std::atof(v) : dflt;
}

// "a,b,c" -> {a,b,c}
inline std::vector<unsigned> env_list(const char *name)
{
    std::vector<unsigned> out;
    std::stringstream ss(env_str(name, ""));
    std::string tok;
    while (std::getline(ss, tok, ','))
        if (!tok.empty())
            out.push_back((unsigned)std::strtoul(tok.c_str(), nullptr, 10));
    return out;
}

inline int host_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// measured copy bandwidth of this pool's B200s (MEASURED_PEAKS.json: hbm_gbs); B200FE_HBM_GBS overrides
inline double hbm_peak_gbs()
{
    return env_double("B200FE_HBM_GBS", 6546.9);
}

// min over `reps` of host time around fn() + device synchronise -- the reference's timing method
template <typename F> double time_min(unsigned reps, F &&fn)
{
    Timer t;
    double best = std::numeric_limits<double>::max();
    for (unsigned r = 0; r < reps; ++r)
    {
        t.start();
        fn();
        CUDA_OK(cudaDeviceSynchronize());
        t.stop();
        best = std::min(best, t.elapsedSeconds());
    }
    return best;
}

template <typename F> double time_min_host(unsigned reps, F &&fn)
{
    Timer t;
    double best = std::numeric_limits<double>::max();
    for (unsigned r = 0; r < reps; ++r)
    {
        t.start();
        fn();
        t.stop();
        best = std::min(best, t.elapsedSeconds());
    }
    return best;
}

// device checksum sum(x^2), the library's replacement of thrust::transform_reduce
template <typename T> struct Checksum
{
    DeviceArray<double> result{1};
    DeviceArray<unsigned char> scratch{b200fe_sumsq_scratch_bytes()};
    // leaves sum(x^2) in *result.get() (device), for the multi-GPU all-reduce
    void launch(const T *x, size_t n)
    {
        if (n == 0)
        {
            CUDA_OK(cudaMemset(result.get(), 0, sizeof(double)));
            return;
        }
        if constexpr (std::is_same<T, double>::value)
            FE_OK(b200fe_sumsq_f64(x, n, result.get(), scratch.get(), nullptr));
        else
            FE_OK(b200fe_sumsq_f32(x, n, result.get(), scratch.get(), nullptr));
    }
    double operator()(const T *x, size_t n)
    {
        launch(x, n);
        double h = 0.0;
        CUDA_OK(cudaMemcpy(&h, result.get(), sizeof(double), cudaMemcpyDeviceToHost));
        return h;
    }
};

// The drivers' repetition loops go through a b200fe_plan: the basis matrices are handed over ONCE per run_test -- where
// the reference uploads them once (benchmark04.cc:890-905, benchmark05.cc:1237-1259) -- and the 40 repetitions of
// every column reuse them, so a repetition is one kernel launch instead of staging launch + kernel.  Same kernels,
// same bits as the per-call entry points (tests/test_plan_gpu.py).  Only for the shapes plans cover (equal nq per
// direction, nm = nq - 1, nq <= 32); B200FE_PLAN=0 keeps every column on the per-call entry points.
class OperatorPlan
{
public:
    OperatorPlan(int dim, bool is_f32, unsigned nq0, unsigned nq1, unsigned nq2, const void *b0, const void *b1,
                 const void *b2)
    {
        const bool regular = nq0 == nq1 && (dim == 2 || nq1 == nq2) && nq0 >= 2u && nq0 <= 32u;
        if (regular && env_long("B200FE_PLAN", 1) != 0)
            FE_OK(b200fe_plan_create(&m_plan, dim, is_f32 ? 1 : 0, nq0, b0, b1, dim == 3 ? b2 : nullptr, nullptr));
    }
    OperatorPlan(const OperatorPlan &)            = delete;
    OperatorPlan &operator=(const OperatorPlan &) = delete;
    ~OperatorPlan()
    {
        if (m_plan)
            b200fe_plan_destroy(m_plan);
    }
    bool active() const { return m_plan != nullptr; }
    void bwdtrans(bool coa, unsigned nelmt, const void *in, void *out) const
    {
        FE_OK(b200fe_plan_bwdtrans(m_plan, coa ? 1 : 0, nelmt, in, out, nullptr));
    }

private:
    b200fe_plan *m_plan = nullptr;
};

template <typename T> double host_sumsq(const std::vector<T> &v)
{
    long double s = 0.0L;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (long long i = 0; i < (long long)v.size(); ++i)
        s += (long double)v[i] * (long double)v[i];
    return (double)s;
}

} // namespace bench

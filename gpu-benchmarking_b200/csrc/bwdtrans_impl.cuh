// bwdtrans_impl.cuh -- launchers + dispatch for one dtype and one dimension.
// Included by bwdtrans_{quad,hex}_{f64,f32}.cu with B200FE_T / B200FE_TAG set.
#pragma once

#include <mutex>

#include "dispatch.h"
#include "sumfac_generic.cuh"
#include "sumfac_rows.cuh"
#include "sumfac_tpe.cuh"

namespace b200fe
{

static BankGuard g_bank;       // per translation unit, like the constant bank itself
static std::mutex g_bank_lock; // fill -> launch -> release is one critical section

inline bool aligned16(const void *p)
{
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

template <typename K> inline int opt_in_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024)
        B200FE_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

// ---- quad ---------------------------------------------------------------------
template <typename T, int NQ, int E, int THREADS>
int launch_quad_rows(unsigned nelmt, const T *in, T *out, cudaStream_t stream)
{
    using C = QuadRows<T, NQ, E, THREADS>;
    static_assert(C::SMEM <= (size_t)kSmemMax, "tile does not fit shared memory");
    auto kernel = bwdtrans_quad_rows_kernel<T, NQ, E, THREADS>;
    int rc      = opt_in_smem(kernel, C::SMEM);
    if (rc)
        return rc;
    const unsigned grid = (nelmt + E - 1) / E;
    const int in_vec    = C::IN_VEC_OK && aligned16(in);
    const int out_vec   = C::OUT_VEC_OK && aligned16(out);
    kernel<<<grid, THREADS, C::SMEM, stream>>>(in, out, nelmt, in_vec, out_vec);
    count_launch();
    t_last_backend = "rows";
    return launch_status();
}

template <typename T, int NQ> int launch_quad_tpe_coa(unsigned nelmt, const T *in, T *out, cudaStream_t stream)
{
    constexpr int THREADS = 128;
    const unsigned grid   = (nelmt + THREADS - 1) / THREADS;
    bwdtrans_quad_tpe_coa_kernel<T, NQ, THREADS><<<grid, THREADS, 0, stream>>>(in, out, nelmt);
    count_launch();
    t_last_backend = "tpe";
    return launch_status();
}

// ---- hex ----------------------------------------------------------------------
template <typename T, int NQ, int E, int THREADS>
int launch_hex_rows(unsigned nelmt, const T *in, T *out, cudaStream_t stream)
{
    using C = HexRows<T, NQ, E, THREADS>;
    static_assert(C::SMEM <= (size_t)kSmemMax, "tile does not fit shared memory");
    auto kernel = bwdtrans_hex_rows_kernel<T, NQ, E, THREADS>;
    int rc      = opt_in_smem(kernel, C::SMEM);
    if (rc)
        return rc;
    const unsigned grid = (nelmt + E - 1) / E;
    const int in_vec    = C::IN_VEC_OK && aligned16(in);
    kernel<<<grid, THREADS, C::SMEM, stream>>>(in, out, nelmt, in_vec);
    count_launch();
    t_last_backend = "rows";
    return launch_status();
}

template <typename T, int NQ> int launch_hex_tpe_coa(unsigned nelmt, const T *in, T *out, cudaStream_t stream)
{
    constexpr int THREADS = 128;
    const unsigned grid   = (nelmt + THREADS - 1) / THREADS;
    bwdtrans_hex_tpe_coa_kernel<T, NQ, THREADS><<<grid, THREADS, 0, stream>>>(in, out, nelmt);
    count_launch();
    t_last_backend = "tpe";
    return launch_status();
}

} // namespace b200fe

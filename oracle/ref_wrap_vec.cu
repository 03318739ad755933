// ref_wrap_vec.cu -- C launch wrappers around the REFERENCE's benchmark01-03 kernels
// (oracle/_ref/b0{1,2,3}_kernels.inc, cut from /root/reference at build time).  TEST
// INFRASTRUCTURE.  Launch shapes follow the reference's run_test()
// (benchmark01.cc:236-249, benchmark02.cc:137-156, benchmark03.cc:248-268).
#include <algorithm>
#include <chrono>
#include <type_traits>

#include <cuda_runtime.h>

#include "utils/cuda_vectors.h" // the reference's own header (found through -I<reference root>)

namespace ref01
{
#include "b01_kernels.inc"
}
namespace ref02
{
#include "b02_kernels.inc"
}
namespace ref03
{
#include "b03_kernels.inc"
}

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, REF_SUF)
using T = REF_T;

extern "C" int FN(ref_set_data)(T *data, unsigned n, cudaStream_t s)
{
    const unsigned blocks = std::min((n + 255u) / 256u, 1024u);
    ref01::set_data<T><<<blocks, 256, 0, s>>>(data, n);
    return (int)cudaGetLastError();
}

// sums must hold min(ceil(n/256), 1024) values; *result = sum x^2.  vl: 0 scalar, 1 vector
extern "C" int FN(ref_l2norm)(T *result, T *sums, T *data, unsigned n, int vl, cudaStream_t s)
{
    const unsigned threads = 256u;
    const unsigned blocks  = std::min((n + threads - 1) / threads, 1024u);
    cudaMemsetAsync(sums, 0, blocks * sizeof(T), s);
    cudaMemsetAsync(result, 0, sizeof(T), s);
    if (vl)
    {
        ref01::l2norm_vl<T, true><<<blocks, threads, 0, s>>>(sums, data, n);
        ref01::reduce_vl<T, true><<<1, blocks, 0, s>>>(result, sums, blocks);
    }
    else
    {
        ref01::l2norm_vl<T, false><<<blocks, threads, 0, s>>>(sums, data, n);
        ref01::reduce_vl<T, false><<<1, blocks, 0, s>>>(result, sums, blocks);
    }
    return (int)cudaGetLastError();
}

extern "C" int FN(ref_add_vector)(T *x, T *y, unsigned n, int vl, cudaStream_t s)
{
    const int threads = 1024;
    const int blocks  = ((n / 8 + threads - 1) / threads);
    if (vl)
        ref02::add_vector<T, true><<<blocks, threads, 0, s>>>(x, y, n);
    else
        ref02::add_vector<T, false><<<blocks, threads, 0, s>>>(x, y, n);
    return (int)cudaGetLastError();
}

extern "C" int FN(ref_matvec)(unsigned N, unsigned M, const T *A, const T *x, T *y, int vl, cudaStream_t s)
{
    const int threads = 256;
    const int blocks  = std::min(M, 65535u);
    if (vl)
        ref03::compute_matvec<T, true><<<blocks, threads, 0, s>>>(N, M, A, x, y);
    else
        ref03::compute_matvec<T, false><<<blocks, threads, 0, s>>>(N, M, A, x, y);
    return (int)cudaGetLastError();
}

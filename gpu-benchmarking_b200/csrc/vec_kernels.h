// vec_kernels.h -- launchers of vec_kernels.cu (internal).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace b200fe
{
template <typename T>
int launch_reduce_partials(T *sums, const T *data, unsigned begin, unsigned end, unsigned slots, bool vl, bool square,
                           cudaStream_t s);
template <typename T> int launch_set_data(T *data, unsigned n, int mode, cudaStream_t s);
template <typename T> int launch_add_vector(T *x, const T *y, unsigned begin, unsigned end, bool vl, cudaStream_t s);
template <typename T> int launch_matvec(unsigned N, unsigned M, const T *A, const T *x, T *y, bool vl, cudaStream_t s);
size_t sumsq_scratch_bytes();
// *result = sum of part[0..n) in a fixed order (second stage of every deterministic reduction)
int launch_sum_final(const double *part, unsigned n, double *result, cudaStream_t s);
template <typename T> int launch_sumsq(const T *x, size_t n, double *result, void *scratch, bool accumulate, cudaStream_t s);
// gemm_form.cu: BwdTrans in its GEMM formulation (intermediates in global memory) and the batched small mat-vec
template <typename T>
int launch_gemm_bwdtrans_quad(unsigned nm0, unsigned nm1, unsigned nq0, unsigned nq1, unsigned nelmt, const T *b0, const T *b1,
                              const T *in, T *wsp, T *out, cudaStream_t s);
template <typename T>
int launch_gemm_bwdtrans_hex(unsigned nm0, unsigned nm1, unsigned nm2, unsigned nq0, unsigned nq1, unsigned nq2, unsigned nelmt,
                             const T *b0, const T *b1, const T *b2, const T *in, T *wsp1, T *wsp2, T *out, cudaStream_t s);
template <typename T>
int launch_matvec_batched(unsigned M, unsigned N, size_t batch, const T *A, size_t strideA, const T *x, size_t stridex, T *y,
                          size_t stridey, cudaStream_t s);
} // namespace b200fe

#pragma once
// cublas_compare.h -- the cuBLAS comparator column of benchmark03-05: the same
// contractions written as (batched) GEMMs on the vendor library, as in the
// reference (benchmark04.cc:750-836, benchmark05.cc:1062-1171).  Library calls
// only; nothing here is part of libb200fe.so.
#include <cublas_v2.h>

#include "bench_common.h"

namespace blascmp
{

#define CUBLAS_OK(expr)                                                                                      \
    do                                                                                                       \
    {                                                                                                        \
        cublasStatus_t s__ = (expr);                                                                         \
        if (s__ != CUBLAS_STATUS_SUCCESS)                                                                    \
            bench::die(#expr, -(int)s__, __FILE__, __LINE__);                                                \
    } while (0)

inline cublasStatus_t gemm(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                           const double *A, int lda, const double *B, int ldb, double *C, int ldc)
{
    const double one = 1.0, zero = 0.0;
    return cublasDgemm(h, ta, tb, m, n, k, &one, A, lda, B, ldb, &zero, C, ldc);
}
inline cublasStatus_t gemm(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                           const float *A, int lda, const float *B, int ldb, float *C, int ldc)
{
    const float one = 1.0f, zero = 0.0f;
    return cublasSgemm(h, ta, tb, m, n, k, &one, A, lda, B, ldb, &zero, C, ldc);
}
inline cublasStatus_t gemm_batched(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                                   const double *A, int lda, long long sa, const double *B, int ldb, long long sb,
                                   double *C, int ldc, long long sc, int batch)
{
    const double one = 1.0, zero = 0.0;
    return cublasDgemmStridedBatched(h, ta, tb, m, n, k, &one, A, lda, sa, B, ldb, sb, &zero, C, ldc, sc, batch);
}
inline cublasStatus_t gemm_batched(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                                   const float *A, int lda, long long sa, const float *B, int ldb, long long sb,
                                   float *C, int ldc, long long sc, int batch)
{
    const float one = 1.0f, zero = 0.0f;
    return cublasSgemmStridedBatched(h, ta, tb, m, n, k, &one, A, lda, sa, B, ldb, sb, &zero, C, ldc, sc, batch);
}
inline cublasStatus_t gemv_t(cublasHandle_t h, int m, int n, const double *A, int lda, const double *x, double *y)
{
    const double one = 1.0, zero = 0.0;
    return cublasDgemv(h, CUBLAS_OP_T, m, n, &one, A, lda, x, 1, &zero, y, 1);
}
inline cublasStatus_t gemv_t(cublasHandle_t h, int m, int n, const float *A, int lda, const float *x, float *y)
{
    const float one = 1.0f, zero = 0.0f;
    return cublasSgemv(h, CUBLAS_OP_T, m, n, &one, A, lda, x, 1, &zero, y, 1);
}

// Basis matrices B[p][i] (i fastest) are column-major nq x nm; element tiles
// in[q][p] (p fastest) are column-major nm0 x nm1.
// quad: W(i, q|e) = B0 * in, one GEMM over all elements; out_e(i, j) = W_e * B1^T, batched over e.
template <typename T>
void bwdtrans_quad(cublasHandle_t h, int nm0, int nm1, int nq0, int nq1, int nelmt, const T *b0, const T *b1,
                   const T *in, T *wsp, T *out)
{
    CUBLAS_OK(gemm(h, CUBLAS_OP_N, CUBLAS_OP_N, nq0, nm1 * nelmt, nm0, b0, nq0, in, nm0, wsp, nq0));
    CUBLAS_OK(gemm_batched(h, CUBLAS_OP_N, CUBLAS_OP_T, nq0, nq1, nm1, wsp, nq0, (long long)nq0 * nm1, b1, nq1, 0LL,
                           out, nq0, (long long)nq0 * nq1, nelmt));
}

// hex, directions 0 -> 1 -> 2:
//   W1(i, rq|e)      = B0 * in                        one GEMM
//   W2_(e,r)(i, j)   = W1_(e,r)(i, q) * B1^T           batched over nelmt*nm2
//   out_e(ji, k)     = W2_e(ji, r) * B2^T              batched over nelmt
template <typename T>
void bwdtrans_hex(cublasHandle_t h, int nm0, int nm1, int nm2, int nq0, int nq1, int nq2, int nelmt, const T *b0,
                  const T *b1, const T *b2, const T *in, T *w1, T *w2, T *out)
{
    CUBLAS_OK(gemm(h, CUBLAS_OP_N, CUBLAS_OP_N, nq0, nm1 * nm2 * nelmt, nm0, b0, nq0, in, nm0, w1, nq0));
    CUBLAS_OK(gemm_batched(h, CUBLAS_OP_N, CUBLAS_OP_T, nq0, nq1, nm1, w1, nq0, (long long)nq0 * nm1, b1, nq1, 0LL, w2,
                           nq0, (long long)nq0 * nq1, nelmt * nm2));
    CUBLAS_OK(gemm_batched(h, CUBLAS_OP_N, CUBLAS_OP_T, nq0 * nq1, nq2, nm2, w2, nq0 * nq1,
                           (long long)nq0 * nq1 * nm2, b2, nq2, 0LL, out, nq0 * nq1, (long long)nq0 * nq1 * nq2,
                           nelmt));
}

} // namespace blascmp

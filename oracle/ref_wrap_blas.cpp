// ref_wrap_blas.cpp -- the reference's cuBLAS comparator column ("cuBLAS", column 5 of the benchmark04/05 logs)
// as C entry points, so bench.py can time it on the same B200 next to libb200fe ("same-box" baseline).
// TEST / MEASUREMENT INFRASTRUCTURE: library calls only, nothing here is part of libb200fe.so.
//
// The GEMM sequences are the reference's, argument for argument:
//   quad  benchmark04/benchmark04.cc:804-820   gemm(N,N) over all elements, then gemmStridedBatched(N,T) per element
//   hex   benchmark05/benchmark05.cc:1128-1153 directions 2 -> 1 -> 0: gemmStridedBatched(N,T), gemm(N,T), gemm(N,T)
// (the result lands in the same out[e][k][j][i] layout; different summation order, so it agrees with the
// kernels to rounding only).  Built by oracle/Makefile into oracle/libref_blas.so (g++, -lcublas -lcudart).
#include <cublas_v2.h>
#include <cuda_runtime_api.h>

namespace
{
cublasHandle_t g_handle[64] = {};

cublasHandle_t handle_for_current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return nullptr;
    if (!g_handle[dev] && cublasCreate(&g_handle[dev]) != CUBLAS_STATUS_SUCCESS)
        g_handle[dev] = nullptr;
    return g_handle[dev];
}

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
inline cublasStatus_t gemm_sb(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                              const double *A, int lda, long long sa, const double *B, int ldb, long long sb, double *C,
                              int ldc, long long sc, int batch)
{
    const double one = 1.0, zero = 0.0;
    return cublasDgemmStridedBatched(h, ta, tb, m, n, k, &one, A, lda, sa, B, ldb, sb, &zero, C, ldc, sc, batch);
}
inline cublasStatus_t gemm_sb(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k,
                              const float *A, int lda, long long sa, const float *B, int ldb, long long sb, float *C,
                              int ldc, long long sc, int batch)
{
    const float one = 1.0f, zero = 0.0f;
    return cublasSgemmStridedBatched(h, ta, tb, m, n, k, &one, A, lda, sa, B, ldb, sb, &zero, C, ldc, sc, batch);
}

// wsp: nelmt*nq0*nm1 values
template <typename T>
int quad(int nq0, int nq1, int nelmt, const T *b0, const T *b1, const T *in, T *wsp, T *out, cudaStream_t s)
{
    cublasHandle_t h = handle_for_current_device();
    if (!h || cublasSetStream(h, s) != CUBLAS_STATUS_SUCCESS)
        return -1;
    const int nm0 = nq0 - 1, nm1 = nq1 - 1;
    cublasStatus_t st = gemm(h, CUBLAS_OP_N, CUBLAS_OP_N, nq0, nm1 * nelmt, nm0, b0, nq0, in, nm0, wsp, nq0);
    if (st == CUBLAS_STATUS_SUCCESS)
        st = gemm_sb(h, CUBLAS_OP_N, CUBLAS_OP_T, nq0, nq1, nm1, wsp, nq0, (long long)nq0 * nm1, b1, nq1, 0LL, out, nq0,
                     (long long)nq0 * nq1, nelmt);
    return -(int)st;
}

// wsp1: nelmt*nq2*nm0*nm1 values, wsp2: nelmt*nq1*nq2*nm0 values
template <typename T>
int hex(int nq0, int nq1, int nq2, int nelmt, const T *b0, const T *b1, const T *b2, const T *in, T *wsp1, T *wsp2,
        T *out, cudaStream_t s)
{
    cublasHandle_t h = handle_for_current_device();
    if (!h || cublasSetStream(h, s) != CUBLAS_STATUS_SUCCESS)
        return -1;
    const int nm0 = nq0 - 1, nm1 = nq1 - 1, nm2 = nq2 - 1;
    cublasStatus_t st = gemm_sb(h, CUBLAS_OP_N, CUBLAS_OP_T, nq2, nm0 * nm1, nm2, b2, nq2, 0LL, in, nm0 * nm1,
                                (long long)nm0 * nm1 * nm2, wsp1, nq2 * nelmt, (long long)nq2, nelmt);
    if (st == CUBLAS_STATUS_SUCCESS)
        st = gemm(h, CUBLAS_OP_N, CUBLAS_OP_T, nq1, nq2 * nelmt * nm0, nm1, b1, nq1, wsp1, nq2 * nelmt * nm0, wsp2, nq1);
    if (st == CUBLAS_STATUS_SUCCESS)
        st = gemm(h, CUBLAS_OP_N, CUBLAS_OP_T, nq0, nq1 * nq2 * nelmt, nm0, b0, nq0, wsp2, nq1 * nq2 * nelmt, out, nq0);
    return -(int)st;
}
} // namespace

extern "C" {
__attribute__((visibility("default"))) int ref_cublas_bwdtrans_quad_f64(int nq0, int nq1, int nelmt, const double *b0,
                                                                        const double *b1, const double *in, double *wsp,
                                                                        double *out, void *stream)
{
    return quad<double>(nq0, nq1, nelmt, b0, b1, in, wsp, out, (cudaStream_t)stream);
}
__attribute__((visibility("default"))) int ref_cublas_bwdtrans_quad_f32(int nq0, int nq1, int nelmt, const float *b0,
                                                                        const float *b1, const float *in, float *wsp,
                                                                        float *out, void *stream)
{
    return quad<float>(nq0, nq1, nelmt, b0, b1, in, wsp, out, (cudaStream_t)stream);
}
__attribute__((visibility("default"))) int ref_cublas_bwdtrans_hex_f64(int nq0, int nq1, int nq2, int nelmt,
                                                                       const double *b0, const double *b1,
                                                                       const double *b2, const double *in, double *wsp1,
                                                                       double *wsp2, double *out, void *stream)
{
    return hex<double>(nq0, nq1, nq2, nelmt, b0, b1, b2, in, wsp1, wsp2, out, (cudaStream_t)stream);
}
__attribute__((visibility("default"))) int ref_cublas_bwdtrans_hex_f32(int nq0, int nq1, int nq2, int nelmt,
                                                                       const float *b0, const float *b1, const float *b2,
                                                                       const float *in, float *wsp1, float *wsp2,
                                                                       float *out, void *stream)
{
    return hex<float>(nq0, nq1, nq2, nelmt, b0, b1, b2, in, wsp1, wsp2, out, (cudaStream_t)stream);
}
}

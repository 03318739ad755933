// dispatch.h -- internal C++ interface between the C ABI (capi.cu) and the
// per-dtype kernel translation units.
#pragma once

#include <cuda_runtime.h>

namespace b200fe
{

enum class Backend : int
{
    Auto    = 0,
    Rows    = 1, // element-batched row contractions through shared memory (element-major)
    Tpe     = 2, // thread per element, registers only (interleaved layout)
    Generic = 3, // run-time sizes, any shape
    Pipe    = 4, // rows + persistent CTAs fed by bulk (TMA) copies through an mbarrier ring
    Nm1     = 6, // nq = 2: scaled broadcast, one thread per 16-byte output chunk, no basis staging
    Lanes   = 7, // lanes = elements, an element's planes / rows split over the warps of a CTA (interleaved layout;
                 // element-major even-nq quads through a bulk-copied slab: "lanes-em")
    Mma     = 5, // FP64 tensor cores (DMMA m8n8k4), one element group per warp, bulk (TMA) fed (quad, even nq)
    Umma    = 8, // FP32 quad nq = 32 on tcgen05.mma kind::tf32 (3xTF32 split, accumulators in TMEM): sumfac_umma.cuh
};

// element-major unless coa; return 0 / cudaError_t / negative B200FE_E*
template <typename T>
int run_bwdtrans_quad(Backend be, bool coa, unsigned nm0, unsigned nm1, unsigned nq0, unsigned nq1, unsigned nelmt,
                      const T *b0, const T *b1, const T *in, T *out, cudaStream_t stream, double *partials = nullptr,
                      unsigned *npartials = nullptr);
template <typename T>
int run_bwdtrans_hex(Backend be, bool coa, unsigned nm0, unsigned nm1, unsigned nm2, unsigned nq0, unsigned nq1,
                     unsigned nq2, unsigned nelmt, const T *b0, const T *b1, const T *b2, const T *in, T *out,
                     cudaStream_t stream, double *partials = nullptr, unsigned *npartials = nullptr);
// partials/npartials: fused checksum.  When the back-end that runs can fuse sum(out^2) into its
// epilogue it writes *npartials (> 0) per-warp partial sums to partials[]; otherwise *npartials stays 0 and the
// caller reduces `out` itself.

// IProductWRTBase (sumfac_iprod.cuh): element-major, nq0 == nq1 (== nq2), nm = nq - 1, nq within the
// per-nq table; w (quadrature metric, one value per point) may be null.  0 / cudaError_t / B200FE_E*
template <typename T>
int run_iproduct_quad(Backend be, unsigned nq, unsigned nelmt, const T *b0, const T *b1, const T *w, const T *in,
                      T *out, cudaStream_t stream);
template <typename T>
int run_iproduct_hex(Backend be, unsigned nq, unsigned nelmt, const T *b0, const T *b1, const T *b2, const T *w,
                     const T *in, T *out, cudaStream_t stream);
// be: Auto = the FP64 tensor-core kernel where one is instantiated, else the row kernel; Rows / Mma force one

} // namespace b200fe

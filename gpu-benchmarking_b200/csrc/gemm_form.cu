// gemm_form.cu -- SURVEY.md 8f-3: (1) BwdTrans in its GEMM formulation -- one contraction per pass with the
// intermediates in GLOBAL memory, the factorisation the reference hands to cuBLAS for its "cuBLAS" column
// (benchmark04.cc:804-820: gemm over all elements + strided-batched gemm; benchmark05.cc:1128-1153) -- on the
// library's own kernels, so the benchmark drivers can fill column 5 without cuBLAS (B200FE_COL5=gemm) and the cost of
// NOT fusing the passes can be read off the same log; (2) a batched small dense mat-vec, benchmark03's operator
// (benchmark03.cc:80-104) for many small matrices at once.
//
// Both accumulate every output in ascending contraction index from 0 with fused multiply-adds: the GEMM formulation is
// bit-identical to the fused BwdTrans entry points (unlike cuBLAS, whose summation order is its own).
#include "common.cuh"
#include "vec_kernels.h"

namespace b200fe
{

constexpr int kGemmThreads = 256;

// pass over contiguous rows:  out[r][i] = sum_p in[r*NMd + p] * B[p*NQd + i],  r < nrows        (direction 0)
// A CTA takes ROWS consecutive rows: their ROWS*nm inputs are one contiguous stretch (coalesced loads into shared
// memory), their ROWS*nq outputs another (coalesced stores).
template <typename T>
__global__ void __launch_bounds__(kGemmThreads)
    gemm_rows_kernel(const T *__restrict__ in, const T *__restrict__ basis, T *__restrict__ out, size_t nrows, unsigned nm,
                     unsigned nq, unsigned rows_per_cta)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_b  = reinterpret_cast<T *>(smem_raw);      // nm * nq
    T *s_in = s_b + nm * nq;                        // rows_per_cta * nm
    for (unsigned t = threadIdx.x; t < nm * nq; t += kGemmThreads)
        s_b[t] = basis[t];
    for (size_t r0 = (size_t)blockIdx.x * rows_per_cta; r0 < nrows; r0 += (size_t)gridDim.x * rows_per_cta)
    {
        const unsigned nr = (unsigned)((nrows - r0 < rows_per_cta) ? nrows - r0 : rows_per_cta);
        __syncthreads();
        for (unsigned t = threadIdx.x; t < nr * nm; t += kGemmThreads)
            s_in[t] = ld_stream(in + r0 * nm + t);
        __syncthreads();
        for (unsigned t = threadIdx.x; t < nr * nq; t += kGemmThreads)
        {
            const unsigned r = t / nq, i = t - r * nq;
            T acc = T(0);
            for (unsigned p = 0; p < nm; ++p)
                acc = fmadd(s_in[r * nm + p], s_b[p * nq + i], acc);
            out[r0 * nq + t] = acc; // the intermediate is read again by the next pass: default caching
        }
    }
}

// pass over a strided index:  out[g][j][w] = sum_q in[g][q][w] * B[q*NQd + j],  g < ngroups, w < width   (directions 1, 2)
// A CTA takes one group and a tile of WT consecutive w: nm x WT inputs into shared memory (coalesced along w), every
// output row j written as a run of WT consecutive values.
template <typename T, bool LAST>
__global__ void __launch_bounds__(kGemmThreads)
    gemm_strided_kernel(const T *__restrict__ in, const T *__restrict__ basis, T *__restrict__ out, size_t ngroups,
                        unsigned nm, unsigned nq, unsigned width, unsigned wt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_b  = reinterpret_cast<T *>(smem_raw); // nm * nq
    T *s_in = s_b + nm * nq;                   // nm * wt
    for (unsigned t = threadIdx.x; t < nm * nq; t += kGemmThreads)
        s_b[t] = basis[t];
    const unsigned tiles_w = (width + wt - 1) / wt;
    const size_t ntiles    = ngroups * tiles_w;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
        const size_t g     = tile / tiles_w;
        const unsigned w0  = (unsigned)(tile - g * tiles_w) * wt;
        const unsigned nw  = (width - w0 < wt) ? width - w0 : wt;
        const T *src       = in + g * (size_t)nm * width + w0;
        T *dst             = out + g * (size_t)nq * width + w0;
        __syncthreads();
        for (unsigned t = threadIdx.x; t < nm * nw; t += kGemmThreads)
        {
            const unsigned q = t / nw, w = t - q * nw;
            s_in[q * wt + w] = src[(size_t)q * width + w];
        }
        __syncthreads();
        for (unsigned t = threadIdx.x; t < nq * nw; t += kGemmThreads)
        {
            const unsigned j = t / nw, w = t - j * nw;
            T acc = T(0);
            for (unsigned q = 0; q < nm; ++q)
                acc = fmadd(s_in[q * wt + w], s_b[q * nq + j], acc);
            if (LAST)
                st_stream(dst + (size_t)j * width + w, acc);
            else
                dst[(size_t)j * width + w] = acc;
        }
    }
}

template <typename T> static int launch_rows(const T *in, const T *b, T *out, size_t nrows, unsigned nm, unsigned nq, cudaStream_t s)
{
    unsigned rows = 2048u / (nm > nq ? nm : nq);
    rows          = rows < 1u ? 1u : rows;
    const size_t smem = ((size_t)nm * nq + (size_t)rows * nm) * sizeof(T);
    if (smem > (size_t)kSmemMax)
        return B200FE_EUNSUPPORTED;
    B200FE_CUDA_TRY(cudaFuncSetAttribute(gemm_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t need   = (nrows + rows - 1) / rows;
    const unsigned grid = (unsigned)(need < 148u * 8u ? need : 148u * 8u);
    gemm_rows_kernel<T><<<grid, kGemmThreads, smem, s>>>(in, b, out, nrows, nm, nq, rows);
    count_launch();
    return launch_status();
}

template <typename T, bool LAST>
static int launch_strided(const T *in, const T *b, T *out, size_t ngroups, unsigned nm, unsigned nq, unsigned width, cudaStream_t s)
{
    unsigned wt = 2048u / (nm > nq ? nm : nq);
    wt          = wt > width ? width : (wt < 1u ? 1u : wt);
    const size_t smem = ((size_t)nm * nq + (size_t)nm * wt) * sizeof(T);
    if (smem > (size_t)kSmemMax)
        return B200FE_EUNSUPPORTED;
    B200FE_CUDA_TRY(cudaFuncSetAttribute(gemm_strided_kernel<T, LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t need   = ngroups * ((width + wt - 1) / wt);
    const unsigned grid = (unsigned)(need < 148u * 8u ? need : 148u * 8u);
    gemm_strided_kernel<T, LAST><<<grid, kGemmThreads, smem, s>>>(in, b, out, ngroups, nm, nq, width, wt);
    count_launch();
    return launch_status();
}

template <typename T>
int launch_gemm_bwdtrans_quad(unsigned nm0, unsigned nm1, unsigned nq0, unsigned nq1, unsigned nelmt, const T *b0, const T *b1,
                              const T *in, T *wsp, T *out, cudaStream_t s)
{
    if (!b0 || !b1 || !in || !wsp || !out || !nm0 || !nm1 || !nq0 || !nq1 || nm0 > 1024u || nm1 > 1024u || nq0 > 1024u ||
        nq1 > 1024u)
        return B200FE_EINVAL;
    if (nelmt == 0)
        return 0;
    // wsp[(e,q)][i] = sum_p in[(e,q)][p] B0[p][i]   -- the reference's gemm(N,N, nq0, nm1*nelmt, nm0)
    int rc = launch_rows<T>(in, b0, wsp, (size_t)nelmt * nm1, nm0, nq0, s);
    if (rc)
        return rc;
    // out[e][j][i] = sum_q wsp[e][q][i] B1[q][j]    -- its gemmStridedBatched(N,T, nq0, nq1, nm1), batch = nelmt
    return launch_strided<T, true>(wsp, b1, out, nelmt, nm1, nq1, nq0, s);
}

template <typename T>
int launch_gemm_bwdtrans_hex(unsigned nm0, unsigned nm1, unsigned nm2, unsigned nq0, unsigned nq1, unsigned nq2, unsigned nelmt,
                             const T *b0, const T *b1, const T *b2, const T *in, T *wsp1, T *wsp2, T *out, cudaStream_t s)
{
    if (!b0 || !b1 || !b2 || !in || !wsp1 || !wsp2 || !out || !nm0 || !nm1 || !nm2 || !nq0 || !nq1 || !nq2 || nm0 > 256u ||
        nm1 > 256u || nm2 > 256u || nq0 > 256u || nq1 > 256u || nq2 > 256u)
        return B200FE_EINVAL;
    if (nelmt == 0)
        return 0;
    // directions 0 -> 1 -> 2 (the order of the fused kernels, so the result is bit-identical to them; the reference's
    // cuBLAS column contracts 2 -> 1 -> 0 and agrees to rounding)
    int rc = launch_rows<T>(in, b0, wsp1, (size_t)nelmt * nm2 * nm1, nm0, nq0, s);                       // [(e,r,q)][i]
    if (rc)
        return rc;
    rc = launch_strided<T, false>(wsp1, b1, wsp2, (size_t)nelmt * nm2, nm1, nq1, nq0, s);               // [(e,r)][j][i]
    if (rc)
        return rc;
    return launch_strided<T, true>(wsp2, b2, out, nelmt, nm2, nq2, nq0 * nq1, s);                        // [e][k][(j,i)]
}

// ---- batched small dense mat-vec -------------------------------------------------------------------------------------
// y_b[i] = sum_j A_b[i*N + j] x_b[j], b < batch; row-major M x N matrices `strideA` values apart (0: one matrix for all),
// vectors stridex / stridey apart.  A CTA stages BPC batch entries (matrices + vectors) in shared memory with coalesced
// loads, then one thread per output row runs the dot product in ascending j.
template <typename T>
__global__ void __launch_bounds__(kGemmThreads)
    matvec_batched_kernel(unsigned M, unsigned N, size_t batch, const T *__restrict__ A, size_t strideA, const T *__restrict__ x,
                          size_t stridex, T *__restrict__ y, size_t stridey, unsigned bpc)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned mn  = M * N, pitch = N + 1; // rows padded by one value: lanes = rows read conflict-free
    T *s_a = reinterpret_cast<T *>(smem_raw);  // bpc (or 1) * M * pitch
    T *s_x = s_a + (strideA ? bpc : 1u) * M * pitch; // bpc * N
    if (!strideA)
        for (unsigned t = threadIdx.x; t < mn; t += kGemmThreads)
            s_a[(t / N) * pitch + t % N] = A[t];
    for (size_t b0 = (size_t)blockIdx.x * bpc; b0 < batch; b0 += (size_t)gridDim.x * bpc)
    {
        const unsigned nb = (unsigned)((batch - b0 < bpc) ? batch - b0 : bpc);
        __syncthreads();
        if (strideA)
            for (unsigned t = threadIdx.x; t < nb * mn; t += kGemmThreads)
            {
                const unsigned b = t / mn, k = t - b * mn;
                s_a[b * M * pitch + (k / N) * pitch + k % N] = ld_stream(A + (b0 + b) * strideA + k);
            }
        for (unsigned t = threadIdx.x; t < nb * N; t += kGemmThreads)
        {
            const unsigned b = t / N, j = t - b * N;
            s_x[t] = x[(b0 + b) * stridex + j];
        }
        __syncthreads();
        for (unsigned t = threadIdx.x; t < nb * M; t += kGemmThreads)
        {
            const unsigned b = t / M, i = t - b * M;
            const T *row = s_a + (strideA ? b * M * pitch : 0u) + i * pitch;
            const T *xv  = s_x + b * N;
            T acc = T(0);
            for (unsigned j = 0; j < N; ++j)
                acc = fmadd(row[j], xv[j], acc);
            y[(b0 + b) * stridey + i] = acc;
        }
    }
}

template <typename T>
int launch_matvec_batched(unsigned M, unsigned N, size_t batch, const T *A, size_t strideA, const T *x, size_t stridex, T *y,
                          size_t stridey, cudaStream_t s)
{
    if (!A || !x || !y || M == 0 || N == 0 || M > 4096u || N > 4096u)
        return B200FE_EINVAL;
    if (batch == 0)
        return 0;
    const size_t per = ((size_t)M * (N + 1) + N) * sizeof(T); // shared memory per staged batch entry
    if (per > (size_t)kSmemMax)
        return B200FE_EUNSUPPORTED; // a matrix that does not fit a CTA's shared memory is benchmark03's job
    size_t bpc = (size_t)(96 * 1024) / per;
    bpc        = bpc < 1 ? 1 : (bpc > 256 ? 256 : bpc);
    if (bpc > batch)
        bpc = batch;
    const size_t smem = strideA ? bpc * per : ((size_t)M * (N + 1) + bpc * N) * sizeof(T);
    B200FE_CUDA_TRY(cudaFuncSetAttribute(matvec_batched_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t need   = (batch + bpc - 1) / bpc;
    const unsigned grid = (unsigned)(need < 148u * 8u ? need : 148u * 8u);
    matvec_batched_kernel<T><<<grid, kGemmThreads, smem, s>>>(M, N, batch, A, strideA, x, stridex, y, stridey, (unsigned)bpc);
    count_launch();
    return launch_status();
}

#define INST(T)                                                                                                        \
    template int launch_gemm_bwdtrans_quad<T>(unsigned, unsigned, unsigned, unsigned, unsigned, const T *, const T *,  \
                                              const T *, T *, T *, cudaStream_t);                                     \
    template int launch_gemm_bwdtrans_hex<T>(unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,     \
                                             const T *, const T *, const T *, const T *, T *, T *, T *, cudaStream_t); \
    template int launch_matvec_batched<T>(unsigned, unsigned, size_t, const T *, size_t, const T *, size_t, T *, size_t, \
                                          cudaStream_t);
INST(double)
INST(float)
#undef INST

} // namespace b200fe

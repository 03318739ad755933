// bwdtrans_quad.cu -- quad BwdTrans dispatch for one dtype.  Compiled twice:
//   -DB200FE_T=double -DB200FE_ROWS_TABLE='"rows_table_2_f64.inc"'
//   -DB200FE_T=float  -DB200FE_ROWS_TABLE='"rows_table_2_f32.inc"'
// The table (tools/gen_rows_table.py + tuner overrides) lists, per nq, the tile
// shape of the rows and pipe back-ends and which of the two the default routing
// prefers.
#include "bwdtrans_impl.cuh"

namespace b200fe
{

using T = B200FE_T;

static int quad_rows_switch(unsigned nq, unsigned nelmt, const T *in, T *out, cudaStream_t s)
{
    switch (nq)
    {
#define ROWS_CASE(NQ, E, TH, R, V)                                                                              \
    case NQ:                                                                                                 \
        return launch_quad_rows<T, NQ, E, TH, R, V>(nelmt, in, out, s);
#define PIPE_CASE(NQ, E, TH, R, V)
#define PREFER(NQ, BE)
#define MMA_CASE(NQ, G, W, MB0, NB1)
#include B200FE_ROWS_TABLE
#undef ROWS_CASE
#undef PIPE_CASE
#undef PREFER
#undef MMA_CASE
    default:
        return B200FE_EUNSUPPORTED;
    }
}

static int quad_pipe_switch(unsigned nq, unsigned nelmt, const T *in, T *out, cudaStream_t s)
{
    switch (nq)
    {
#define ROWS_CASE(NQ, E, TH, R, V)
#define PIPE_CASE(NQ, E, TH, R, V)                                                                              \
    case NQ:                                                                                                 \
        return launch_quad_pipe<T, NQ, E, TH, R, V>(nelmt, in, out, s);
#define PREFER(NQ, BE)
#define MMA_CASE(NQ, G, W, MB0, NB1)
#include B200FE_ROWS_TABLE
#undef ROWS_CASE
#undef PIPE_CASE
#undef PREFER
#undef MMA_CASE
    default:
        return B200FE_EUNSUPPORTED;
    }
}

// tensor-core variants: DMMA for T = double (sumfac_mma.cuh), 3xTF32 for T = float (sumfac_mma32.cuh);
// launch_quad_mma is overloaded on the pointer type, the table of each dtype lists its own MMA_CASE lines
static int quad_mma_switch(unsigned nq, unsigned nelmt, const T *b0, const T *b1, const T *in, T *out, cudaStream_t s,
                           double *partials, unsigned *npartials)
{
    switch (nq)
    {
#define ROWS_CASE(NQ, E, TH, R, V)
#define PIPE_CASE(NQ, E, TH, R, V)
#define PREFER(NQ, BE)
#define MMA_CASE(NQ, G, W, MB0, NB1)                                                                         \
    case NQ:                                                                                                 \
        return launch_quad_mma<NQ, G, W, MB0, NB1>(nelmt, b0, b1, in, out, s, partials, npartials);
#include B200FE_ROWS_TABLE
#undef ROWS_CASE
#undef PIPE_CASE
#undef PREFER
#undef MMA_CASE
    default:
        return B200FE_EUNSUPPORTED;
    }
}

// interleaved layout through the rows passes (tile shape of the element-major rows entry)
static int quad_rowscoa_switch(unsigned nq, unsigned nelmt, const T *in, T *out, cudaStream_t s)
{
    switch (nq)
    {
#define ROWS_CASE(NQ, E, TH, R, V)                                                                              \
    case NQ:                                                                                                 \
        return launch_quad_rowscoa<T, NQ, E, TH, R, V>(nelmt, in, out, s);
#define PIPE_CASE(NQ, E, TH, R, V)
#define PREFER(NQ, BE)
#define MMA_CASE(NQ, G, W, MB0, NB1)
#include B200FE_ROWS_TABLE
#undef ROWS_CASE
#undef PIPE_CASE
#undef PREFER
#undef MMA_CASE
    default:
        return B200FE_EUNSUPPORTED;
    }
}

// what the table offers for this nq: bit 0 rows, bit 1 pipe, bit 2 mma; *preferred = default routing
static int quad_table_lookup(unsigned nq, Backend *preferred)
{
    int have   = 0;
    *preferred = Backend::Generic;
    switch (nq)
    {
#define ROWS_CASE(NQ, E, TH, R, V)                                                                              \
    case NQ:                                                                                                 \
        have |= 1;                                                                                           \
        break;
#define PIPE_CASE(NQ, E, TH, R, V)
#define PREFER(NQ, BE)
#define MMA_CASE(NQ, G, W, MB0, NB1)
#include B200FE_ROWS_TABLE
#undef ROWS_CASE
#undef PIPE_CASE
#undef PREFER
#undef MMA_CASE
    default:
        break;
    }
    switch (nq)
    {
#define ROWS_CASE(NQ, E, TH, R, V)
#define PIPE_CASE(NQ, E, TH, R, V)                                                                              \
    case NQ:                                                                                                 \
        have |= 2;                                                                                           \
        break;
#define PREFER(NQ, BE)
#define MMA_CASE(NQ, G, W, MB0, NB1)
#include B200FE_ROWS_TABLE
#undef ROWS_CASE
#undef PIPE_CASE
#undef PREFER
#undef MMA_CASE
    default:
        break;
    }
    switch (nq)
    {
#define ROWS_CASE(NQ, E, TH, R, V)
#define PIPE_CASE(NQ, E, TH, R, V)
#define PREFER(NQ, BE)
#define MMA_CASE(NQ, G, W, MB0, NB1)                                                                         \
    case NQ:                                                                                                 \
        have |= 4;                                                                                           \
        break;
#include B200FE_ROWS_TABLE
#undef ROWS_CASE
#undef PIPE_CASE
#undef PREFER
#undef MMA_CASE
    default:
        break;
    }
    switch (nq)
    {
#define ROWS_CASE(NQ, E, TH, R, V)
#define PIPE_CASE(NQ, E, TH, R, V)
#define PREFER(NQ, BE)                                                                                       \
    case NQ:                                                                                                 \
        *preferred = Backend::BE;                                                                            \
        break;
#define MMA_CASE(NQ, G, W, MB0, NB1)
#include B200FE_ROWS_TABLE
#undef ROWS_CASE
#undef PIPE_CASE
#undef PREFER
#undef MMA_CASE
    default:
        break;
    }
    return have;
}

// interleaved layout, lanes back-end (sumfac_lanes.cuh): nq = 4 .. 16 and 32.  A CTA takes a whole interleave group
// (EL = 32) except where the block would exceed 1024 threads or, FP64 nq = 16, two half groups per SM measured better
// (tools/tune/lanes_probe.cu, profiles/r01_lanes_probe.csv).
// Against the thread-per-element kernel (tools/coa_compare.py, profiles/r01_coa_compare.csv): FP64 lanes wins from
// nq = 4 (0.96-1.01 of the measured copy bandwidth against 0.67-0.97), FP32 from nq = 7 (small CTAs below that).
constexpr unsigned kQuadLanesMinNq = sizeof(T) == 8 ? 4 : 7;
static bool quad_has_lanes(unsigned nq)
{
    return (nq >= 4 && nq <= 16) || nq == 32;
}
static int quad_lanes_switch(unsigned nq, unsigned nelmt, const T *in, T *out, cudaStream_t s)
{
    constexpr bool D = sizeof(T) == 8;
    switch (nq)
    {
#define LANES_CASE(NQ, EL)                                                                                   \
    case NQ:                                                                                                 \
        return launch_quad_lanes<T, NQ, EL>(nelmt, in, out, s);
        LANES_CASE(4, 32)
        LANES_CASE(5, 32)
        LANES_CASE(6, 32)
        LANES_CASE(7, 32)
        LANES_CASE(8, 32)
        LANES_CASE(9, 32)
        LANES_CASE(10, 32)
        LANES_CASE(11, 32)
        LANES_CASE(12, 32)
        LANES_CASE(13, 32)
        LANES_CASE(14, 32)
        LANES_CASE(15, 32)
        LANES_CASE(16, (D ? 16 : 32))
        LANES_CASE(32, 16)
#undef LANES_CASE
    default:
        return B200FE_EUNSUPPORTED;
    }
}

// element-major lanes kernel (sumfac_lanes.cuh, "lanes-em"): even nq where it measured faster than the table's choice
// at 64 Mi points (tools/tune/lanesem_probe.cu, profiles/r01_lanesem_probe.csv; fraction of the HBM roofline).
// TPC = tiles per CTA (all bulk copies issued up front): pays for the small nq, whose tiles are only 3-14 KB.
//   FP64  nq   4     6     8     10    12    14    16      FP32  nq   4     6     8     10    12    14    16
//   EL / TPC  32/4  16/2  32/1  16/1  16/1   4/1   4/1           EL  64/8  16/8  32/4   8/4  16/1   8/1   8/1
//   lanes-em  0.98  0.98  1.00  0.99  0.99  0.98  0.97               0.90  0.89  0.91  0.89  0.92  0.89  0.93
//   before    0.94  0.89  0.93  0.91  0.96  0.94  0.92               0.88  0.85  0.89  0.85  0.85  0.66  0.82
static bool quad_has_lanesem(unsigned nq)
{
    if (sizeof(T) == 8)
        return nq % 2 == 0 && nq >= 4 && nq <= 16;
    return nq % 2 == 0 && nq >= 4 && nq <= 16;
}
static int quad_lanesem_switch(unsigned nq, unsigned nelmt, const T *in, T *out, cudaStream_t s, double *partials,
                               unsigned *npartials)
{
    constexpr bool D = sizeof(T) == 8;
    switch (nq)
    {
    case 4:
        return launch_quad_lanesem<T, 4, (D ? 32 : 64), (D ? 4 : 8)>(nelmt, in, out, s, partials, npartials);
    case 6:
        return launch_quad_lanesem<T, 6, 16, (D ? 2 : 8)>(nelmt, in, out, s, partials, npartials);
    case 8:
        return launch_quad_lanesem<T, 8, 32, (D ? 1 : 4)>(nelmt, in, out, s, partials, npartials);
    case 10:
        return launch_quad_lanesem<T, 10, (D ? 16 : 8), (D ? 1 : 4)>(nelmt, in, out, s, partials, npartials);
    case 12:
        return launch_quad_lanesem<T, 12, 16>(nelmt, in, out, s, partials, npartials);
    case 14:
        return launch_quad_lanesem<T, 14, (D ? 4 : 8)>(nelmt, in, out, s, partials, npartials);
    case 16:
        return launch_quad_lanesem<T, 16, (D ? 4 : 8)>(nelmt, in, out, s, partials, npartials);
    default:
        break;
    }
    return B200FE_EUNSUPPORTED;
}

// registers hold nm^2 + nm values per thread

static int quad_tpe_switch(unsigned nq, unsigned nelmt, const T *in, T *out, cudaStream_t s)
{
    switch (nq)
    {
#define TPE_CASE(NQ)                                                                                         \
    case NQ:                                                                                                 \
        return launch_quad_tpe_coa<T, NQ>(nelmt, in, out, s);
        TPE_CASE(2)
        TPE_CASE(3)
        TPE_CASE(4)
        TPE_CASE(5)
        TPE_CASE(6)
        TPE_CASE(7)
        TPE_CASE(8)
        TPE_CASE(9)
        TPE_CASE(10)
#undef TPE_CASE
    default:
        return B200FE_EUNSUPPORTED;
    }
}

// interleaved layout on the tensor cores with M = elements (sumfac_coamma.cuh): the one compute-bound case.
// FP64 on DMMA, bit-identical; FP32 on 3xTF32 mma.sync, held to the component-wise 1e-5 bound of include/b200fe.h
static bool quad_has_coamma(unsigned nq)
{
    return nq == 32;
}

template <>
int run_bwdtrans_quad<T>(Backend be, bool coa, unsigned nm0, unsigned nm1, unsigned nq0, unsigned nq1, unsigned nelmt,
                        const T *b0, const T *b1, const T *in, T *out, cudaStream_t stream, double *partials,
                        unsigned *npartials)
{
    if (npartials)
        *npartials = 0;
    const bool regular = (nq0 == nq1) && (nm0 + 1 == nq0) && (nm1 + 1 == nq1) && nq0 >= 2 && nq0 <= 32;
    Backend preferred  = Backend::Generic;
    const int have     = regular ? quad_table_lookup(nq0, &preferred) : 0;
    if (be == Backend::Auto)
    {
        if (!regular)
            be = Backend::Generic;
        else if (coa)
            be = nq0 < kQuadLanesMinNq ? Backend::Tpe
                 : (quad_has_coamma(nq0) && aligned16(in)) ? Backend::Mma // nq = 32: FP64 0.55 against 0.29 (lanes), FP32 0.48 against 0.33
                 : quad_has_lanes(nq0) ? Backend::Lanes
                                       : ((have & 1) ? Backend::Rows : Backend::Generic);
        else if (nq0 == 2 && sizeof(T) == 4)
            be = Backend::Nm1; // measured: 0.81 vs 0.63 (pipe) for FP32; FP64 and hex stay on the table's choice
        else if (quad_has_lanesem(nq0) && aligned16(in))
            be = Backend::Lanes;
        else if (nq0 == 32 && sizeof(T) == 4 && aligned16(in))
            be = Backend::Umma; // tcgen05 kind::tf32 (sumfac_umma.cuh); a misaligned `in` stays on the warp-level path
        else
            be = preferred;
        // the bulk-copy ring needs a 16-byte aligned slab; otherwise take the plain-load twin
        if (be == Backend::Pipe && (!(have & 2) || !aligned16(in)))
            be = (have & 1) ? Backend::Rows : Backend::Generic;
        if (be == Backend::Mma && !coa && !(have & 4))
            be = (have & 1) ? Backend::Rows : Backend::Generic;
        if (be == Backend::Rows && !(have & 1))
            be = Backend::Generic;
    }
    if (be == Backend::Generic)
    {
        t_last_backend = "generic";
        return launch_quad_generic<T>(nm0, nm1, nq0, nq1, nelmt, b0, b1, in, out, coa, stream);
    }
    if (!regular || ((be == Backend::Pipe || be == Backend::Nm1 || be == Backend::Umma) && coa) ||
        (be == Backend::Tpe && !coa) || (be == Backend::Rows && !(have & 1)))
        return B200FE_EUNSUPPORTED;
    if (be == Backend::Mma && coa) // DMMA with M = elements (sumfac_coamma.cuh): basis from global memory, no bank
        return (quad_has_coamma(nq0) && aligned16(in)) ? launch_quad_coamma<32, 4>(nelmt, b0, b1, in, out, stream)
                                                        : B200FE_EUNSUPPORTED;
    if (be == Backend::Lanes && !coa && (!quad_has_lanesem(nq0) || !aligned16(in)))
        return B200FE_EUNSUPPORTED; // the bulk copy of the slab needs a 16-byte aligned `in`
    if (be == Backend::Pipe && (!(have & 2) || !aligned16(in)))
        return B200FE_EUNSUPPORTED;
    if (be == Backend::Nm1)
        return (nq0 == 2 && !coa) ? launch_nm1<T, 2>(nelmt, b0, b1, b1, in, out, stream) : B200FE_EUNSUPPORTED;
    if (be == Backend::Umma)
        return (nq0 == 32 && sizeof(T) == 4 && !coa && aligned16(in))
                   ? launch_quad_umma<T>(nelmt, b0, b1, in, out, stream, partials, npartials)
                   : B200FE_EUNSUPPORTED;
    if (be == Backend::Mma) // reads the basis matrices from global memory: no constant bank, no lock
        return (have & 4) ? quad_mma_switch(nq0, nelmt, b0, b1, in, out, stream, partials, npartials) : B200FE_EUNSUPPORTED;

    std::lock_guard<std::mutex> lock(bank_lock_of_current_device());
    const T *bases[2]   = {b0, b1};
    int rc              = fill_basis_bank<T>(g_bank, 2, bases, (int)nm0, (int)nq0, false, stream);
    if (rc)
        return rc;
    if (be == Backend::Rows)
        rc = coa ? quad_rowscoa_switch(nq0, nelmt, in, out, stream) : quad_rows_switch(nq0, nelmt, in, out, stream);
    else if (be == Backend::Pipe)
        rc = quad_pipe_switch(nq0, nelmt, in, out, stream);
    else if (be == Backend::Lanes)
        rc = coa ? quad_lanes_switch(nq0, nelmt, in, out, stream) : quad_lanesem_switch(nq0, nelmt, in, out, stream, partials, npartials);
    else
        rc = quad_tpe_switch(nq0, nelmt, in, out, stream);
    // the fill is enqueued: record the bank's event on the error path too, or another stream's next fill could
    // overlap it
    const int rel = release_basis_bank(g_bank, stream);
    return rc ? rc : rel;
}

// ---- IProductWRTBase ---------------------------------------------------------------
static int quad_iprod_switch(unsigned nq, unsigned nelmt, const T *in, const T *w, T *out, cudaStream_t s)
{
    switch (nq)
    {
#define ROWS_CASE(NQ, E, TH, R, V)                                                                              \
    case NQ:                                                                                                 \
        return launch_quad_iprod<T, NQ, E, TH, R>(nelmt, in, w, out, s);
#define PIPE_CASE(NQ, E, TH, R, V)
#define PREFER(NQ, BE)
#define MMA_CASE(NQ, G, W, MB0, NB1)
#include B200FE_ROWS_TABLE
#undef ROWS_CASE
#undef PIPE_CASE
#undef PREFER
#undef MMA_CASE
    default:
        return B200FE_EUNSUPPORTED;
    }
}

// FP64 tensor-core variant: (G elements per warp, warps per CTA, MB, NB) per nq
static int quad_iprod_mma_switch(unsigned nq, unsigned nelmt, const T *b0, const T *b1, const T *in, const T *w, T *out,
                                 cudaStream_t s)
{
    {
        switch (nq)
        {
#define IPM(NQ, G, W, MB, NB)                                                                                \
    case NQ:                                                                                                 \
        return launch_quad_iprod_mma<NQ, G, W, MB, NB>(nelmt, b0, b1, in, w, out, s);
            // (G, WARPS) from a sweep at 64 Mi points (tools/ipm_probe.py): fraction of the roofline, row kernel in ()
            IPM(8, 8, 4, 4, 4)  // 0.89 (0.65)
            IPM(12, 2, 4, 4, 4) // 0.63 (0.57)
            IPM(14, 2, 4, 4, 4) // 0.64 (0.57)
            IPM(16, 2, 8, 4, 4) // 0.70 (0.63)
            IPM(32, 1, 2, 4, 4) // 0.40 (0.23)
#undef IPM
        default:
            break;
        }
    }
    return B200FE_EUNSUPPORTED;
}

// lanes-style kernel (sumfac_iprod_lanes.cuh): even nq, 16-byte aligned in / w.  Elements per CTA from
// tools/ipl_probe.py at 64 Mi points (profiles/r01_ipl_probe.csv); fraction of the roofline unweighted / weighted,
// row or tensor-core kernel in ():
//   FP64 nq   4            6            8            10           12           14           16
//   EL        32           8            4            8            8            16           2
//             0.96 / 1.03  0.97 / 1.05  0.92 / 1.05  0.97 / 1.02  0.98 / 1.02  0.93 / 0.90  0.81 / 0.75
//            (0.57 / 0.85)(0.54 / 0.73)(0.90 / 1.00)(0.63 / 0.88)(0.62 / 0.88)(0.64 / 0.79)(0.69 / 0.73)
//   FP32 EL   32           16           8            8            8            8            4
//             0.79 / 1.00  0.85 / 1.01  0.83 / 1.01  0.86 / 0.99  0.90 / 0.99  0.82 / 0.97  0.80 / 0.93
//            (0.56 / 0.86)(0.47 / 0.65)(0.53 / 0.72)(0.46 / 0.60)(0.56 / 0.75)(0.33 / 0.46)(0.42 / 0.63)
static bool quad_has_iprod_lanes(unsigned nq, bool)
{
    return nq % 2 == 0 && nq >= 4 && nq <= 16;
}
static int quad_iprod_lanes_switch(unsigned nq, unsigned nelmt, const T *in, const T *w, T *out, cudaStream_t s)
{
    constexpr bool D = sizeof(T) == 8;
    switch (nq)
    {
#define IPL(NQ, EL)                                                                                          \
    case NQ:                                                                                                 \
        return launch_quad_iprod_lanes<T, NQ, EL>(nelmt, in, w, out, s);
        IPL(4, 32)
        IPL(6, (D ? 8 : 16))
        IPL(8, (D ? 4 : 8))
        IPL(10, 8)
        IPL(12, 8)
        IPL(14, (D ? 16 : 8))
        IPL(16, (D ? 2 : 4))
#undef IPL
    default:
        return B200FE_EUNSUPPORTED;
    }
}

template <>
int run_iproduct_quad<T>(Backend be, unsigned nq, unsigned nelmt, const T *b0, const T *b1, const T *w, const T *in,
                         T *out, cudaStream_t stream)
{
    const bool aligned = aligned16(in) && (!w || aligned16(w)); // rows are fetched with 8- / 16-byte loads
    if (be == Backend::Lanes && !(quad_has_iprod_lanes(nq, false) && aligned))
        return B200FE_EUNSUPPORTED;
    const bool lanes = be == Backend::Lanes || (be == Backend::Auto && quad_has_iprod_lanes(nq, w != nullptr) && aligned);
    if (!lanes && (be == Backend::Auto || be == Backend::Mma))
    {
        const int rc = quad_iprod_mma_switch(nq, nelmt, b0, b1, in, w, out, stream);
        if (rc != B200FE_EUNSUPPORTED || be == Backend::Mma)
            return rc;
    }
    else if (!lanes && be != Backend::Rows)
        return B200FE_EUNSUPPORTED;
    std::lock_guard<std::mutex> lock(bank_lock_of_current_device());
    const T *bases[2]   = {b0, b1};
    int rc = fill_basis_bank<T>(g_bank, 2, bases, (int)nq - 1, (int)nq, true, stream); // transposed
    if (rc)
        return rc;
    rc = lanes ? quad_iprod_lanes_switch(nq, nelmt, in, w, out, stream) : quad_iprod_switch(nq, nelmt, in, w, out, stream);
    // the fill is enqueued: record the bank's event on the error path too, or another stream's next fill could
    // overlap it
    const int rel = release_basis_bank(g_bank, stream);
    return rc ? rc : rel;
}

} // namespace b200fe

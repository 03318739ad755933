#pragma once
// cpu_reference.h -- the host-CPU comparator columns of the benchmark drivers.
//
// The reference's only backend-portable code is its Kokkos variants; run on a
// Kokkos host backend they are plain per-element loop nests under a parallel
// for.  These are those loop nests in C++17 + OpenMP (element-major nest:
// benchmark04.cc:493-518 / benchmark05.cc:695-741; interleaved nest:
// benchmark04.cc:536-567 / benchmark05.cc:759-815; two-pass sum-factorised
// form: benchmark04.cc:260-294 / benchmark05.cc:361-423).  They fill the log
// columns the reference used for Kokkos, so every GPU number sits next to the
// same operator on the box's own host cores.  They are benchmark columns, not a
// fallback: nothing in libb200fe.so calls them.
#include <cstddef>
#include <vector>

namespace cpuref
{

enum class Layout
{
    ElementMajor,
    Interleaved
};

inline size_t at(Layout l, size_t e, size_t k, size_t len)
{
    return l == Layout::ElementMajor ? e * len + k : (e / 32) * 32 * len + 32 * k + (e % 32);
}

// quad, per-column nest: for each i contract p into a scratch row, then q for every j
template <typename T>
void bwdtrans_quad_nest(Layout l, unsigned nm0, unsigned nm1, unsigned nq0, unsigned nq1, size_t nelmt,
                        const T *b0, const T *b1, const T *in, T *out)
{
    const size_t nmTot = (size_t)nm0 * nm1, nqTot = (size_t)nq0 * nq1;
#pragma omp parallel for schedule(static)
    for (long long ee = 0; ee < (long long)nelmt; ++ee)
    {
        const size_t e = (size_t)ee;
        std::vector<T> row(nm1);
        for (unsigned i = 0; i < nq0; ++i)
        {
            for (unsigned q = 0; q < nm1; ++q)
            {
                T acc = 0;
                for (unsigned p = 0; p < nm0; ++p)
                    acc += in[at(l, e, (size_t)q * nm0 + p, nmTot)] * b0[p * nq0 + i];
                row[q] = acc;
            }
            for (unsigned j = 0; j < nq1; ++j)
            {
                T acc = 0;
                for (unsigned q = 0; q < nm1; ++q)
                    acc += row[q] * b1[q * nq1 + j];
                out[at(l, e, (size_t)j * nq0 + i, nqTot)] = acc;
            }
        }
    }
}

// quad, two-pass sum factorisation with a per-thread intermediate tile [i][q]
template <typename T>
void bwdtrans_quad_twopass(unsigned nm0, unsigned nm1, unsigned nq0, unsigned nq1, size_t nelmt, const T *b0,
                           const T *b1, const T *in, T *out)
{
    const size_t nmTot = (size_t)nm0 * nm1, nqTot = (size_t)nq0 * nq1;
#pragma omp parallel for schedule(static)
    for (long long ee = 0; ee < (long long)nelmt; ++ee)
    {
        const T *src = in + (size_t)ee * nmTot;
        T *dst       = out + (size_t)ee * nqTot;
        std::vector<T> tile((size_t)nq0 * nm1);
        for (unsigned i = 0; i < nq0; ++i)
            for (unsigned q = 0; q < nm1; ++q)
            {
                T acc = 0;
                for (unsigned p = 0; p < nm0; ++p)
                    acc += src[q * nm0 + p] * b0[p * nq0 + i];
                tile[(size_t)i * nm1 + q] = acc;
            }
        for (unsigned j = 0; j < nq1; ++j)
            for (unsigned i = 0; i < nq0; ++i)
            {
                T acc = 0;
                for (unsigned q = 0; q < nm1; ++q)
                    acc += tile[(size_t)i * nm1 + q] * b1[q * nq1 + j];
                dst[j * nq0 + i] = acc;
            }
    }
}

// hex, nested nest (i, then j, then k)
template <typename T>
void bwdtrans_hex_nest(Layout l, unsigned nm0, unsigned nm1, unsigned nm2, unsigned nq0, unsigned nq1, unsigned nq2,
                       size_t nelmt, const T *b0, const T *b1, const T *b2, const T *in, T *out)
{
    const size_t nmTot = (size_t)nm0 * nm1 * nm2, nqTot = (size_t)nq0 * nq1 * nq2;
#pragma omp parallel for schedule(static)
    for (long long ee = 0; ee < (long long)nelmt; ++ee)
    {
        const size_t e = (size_t)ee;
        std::vector<T> plane((size_t)nm1 * nm2), line(nm2);
        for (unsigned i = 0; i < nq0; ++i)
        {
            for (unsigned rq = 0; rq < nm1 * nm2; ++rq)
            {
                T acc = 0;
                for (unsigned p = 0; p < nm0; ++p)
                    acc += in[at(l, e, (size_t)rq * nm0 + p, nmTot)] * b0[p * nq0 + i];
                plane[rq] = acc;
            }
            for (unsigned j = 0; j < nq1; ++j)
            {
                for (unsigned r = 0; r < nm2; ++r)
                {
                    T acc = 0;
                    for (unsigned q = 0; q < nm1; ++q)
                        acc += plane[(size_t)r * nm1 + q] * b1[q * nq1 + j];
                    line[r] = acc;
                }
                for (unsigned k = 0; k < nq2; ++k)
                {
                    T acc = 0;
                    for (unsigned r = 0; r < nm2; ++r)
                        acc += line[r] * b2[r * nq2 + k];
                    out[at(l, e, ((size_t)k * nq1 + j) * nq0 + i, nqTot)] = acc;
                }
            }
        }
    }
}

// hex, three-pass sum factorisation with per-thread intermediates [i][r][q] and [j][i][r]
template <typename T>
void bwdtrans_hex_threepass(unsigned nm0, unsigned nm1, unsigned nm2, unsigned nq0, unsigned nq1, unsigned nq2,
                            size_t nelmt, const T *b0, const T *b1, const T *b2, const T *in, T *out)
{
    const size_t nmTot = (size_t)nm0 * nm1 * nm2, nqTot = (size_t)nq0 * nq1 * nq2;
#pragma omp parallel for schedule(static)
    for (long long ee = 0; ee < (long long)nelmt; ++ee)
    {
        const T *src = in + (size_t)ee * nmTot;
        T *dst       = out + (size_t)ee * nqTot;
        std::vector<T> t1((size_t)nq0 * nm1 * nm2), t2((size_t)nq0 * nq1 * nm2);
        for (unsigned i = 0; i < nq0; ++i)
            for (unsigned rq = 0; rq < nm1 * nm2; ++rq)
            {
                T acc = 0;
                for (unsigned p = 0; p < nm0; ++p)
                    acc += src[(size_t)rq * nm0 + p] * b0[p * nq0 + i];
                t1[(size_t)i * nm1 * nm2 + rq] = acc;
            }
        for (unsigned j = 0; j < nq1; ++j)
            for (unsigned i = 0; i < nq0; ++i)
                for (unsigned r = 0; r < nm2; ++r)
                {
                    T acc = 0;
                    for (unsigned q = 0; q < nm1; ++q)
                        acc += t1[((size_t)i * nm2 + r) * nm1 + q] * b1[q * nq1 + j];
                    t2[((size_t)j * nq0 + i) * nm2 + r] = acc;
                }
        for (unsigned k = 0; k < nq2; ++k)
            for (unsigned ji = 0; ji < nq0 * nq1; ++ji)
            {
                T acc = 0;
                for (unsigned r = 0; r < nm2; ++r)
                    acc += t2[(size_t)ji * nm2 + r] * b2[r * nq2 + k];
                dst[(size_t)k * nq0 * nq1 + ji] = acc;
            }
    }
}

} // namespace cpuref

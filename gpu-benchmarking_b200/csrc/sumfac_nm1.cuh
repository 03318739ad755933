// sumfac_nm1.cuh -- "nm1" back-end: nq = 2 (one mode per direction), element-major.
//
// With nm = 1 the operator degenerates to a scaled broadcast,
//     out[e][k][j][i] = ((in[e] * B0[i]) * B1[j]) * B2[k],
// 4 (quad) or 8 (hex) outputs per input value: a pure stream with a 1:4 / 1:8 read:write ratio.  One
// thread produces one 16-byte chunk of `out` (coalesced 512-byte warp stores) from one (broadcast) load;
// the six basis values come straight from global memory, so there is no constant-bank fill launch in front
// of a kernel that only runs for 50-100 us.  Every product goes through fma(a, b, +0) -- the reference's
// `tmp = 0; tmp += a*b` as nvcc contracts it (benchmark04.cc:55-59) -- so signed zeros match bit for bit.
#pragma once

#include "common.cuh"

namespace b200fe
{

template <typename T, int DIM>
__global__ void __launch_bounds__(256)
    bwdtrans_nm1_kernel(const T *__restrict__ b0, const T *__restrict__ b1, const T *__restrict__ b2,
                        const T *__restrict__ in, T *__restrict__ out, size_t nchunks, int out_vec)
{
    using V              = typename Vec16<T>::type;
    constexpr int W      = Vec16<T>::W;      // outputs per chunk
    constexpr int NOUT   = 1 << DIM;         // outputs per element
    constexpr int CHUNKS = NOUT / W;         // chunks per element (1, 2 or 4)
    const T B0[2] = {__ldg(b0), __ldg(b0 + 1)}, B1[2] = {__ldg(b1), __ldg(b1 + 1)};
    T B2[2] = {T(1), T(1)};
    if (DIM == 3)
    {
        B2[0] = __ldg(b2);
        B2[1] = __ldg(b2 + 1);
    }
    auto produce = [&](size_t c, T x) {
        const int part = (int)(c % CHUNKS);
        T v[W];
#pragma unroll
        for (int w = 0; w < W; ++w)
        {
            const int o = part * W + w;
            T t         = fmadd(x, B0[o & 1], T(0));
            t           = fmadd(t, B1[(o >> 1) & 1], T(0));
            if (DIM == 3)
                t = fmadd(t, B2[(o >> 2) & 1], T(0));
            v[w] = t;
        }
        if (out_vec)
        {
            V pack;
            if constexpr (W == 2)
                pack = make_double2(v[0], v[1]);
            else
                pack = make_float4(v[0], v[1], v[2], v[3]);
            st_stream(reinterpret_cast<V *>(out) + c, pack);
        }
        else
        {
#pragma unroll
            for (int w = 0; w < W; ++w)
                st_stream(out + c * W + w, v[w]);
        }
    };
    // UNROLL independent loads in flight per thread: with one, a thread moves 16 bytes per memory round trip and
    // the kernel is latency-bound at ~0.77 of the roofline
    constexpr int UNROLL = 8;
    const size_t stride  = (size_t)gridDim.x * blockDim.x;
    size_t c             = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; c + (UNROLL - 1) * stride < nchunks; c += UNROLL * stride)
    {
        T x[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            x[u] = ld_stream(in + (c + u * stride) / CHUNKS);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            produce(c + u * stride, x[u]);
    }
    for (; c < nchunks; c += stride)
        produce(c, ld_stream(in + c / CHUNKS));
}

template <typename T, int DIM>
int launch_nm1(unsigned nelmt, const T *b0, const T *b1, const T *b2, const T *in, T *out, cudaStream_t stream)
{
    constexpr int CHUNKS = (1 << DIM) / Vec16<T>::W;
    const size_t nchunks = (size_t)nelmt * CHUNKS;
    const size_t want    = (nchunks + 255) / 256;
    const unsigned cap   = 148u * 8u * 2u; // two waves of 8 resident CTAs per SM, grid-stride beyond
    const unsigned grid  = (unsigned)(want < cap ? want : cap);
    const int out_vec    = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    bwdtrans_nm1_kernel<T, DIM><<<grid, 256, 0, stream>>>(b0, b1, b2, in, out, nchunks, out_vec);
    count_launch();
    t_last_backend = "nm1";
    return launch_status();
}

} // namespace b200fe

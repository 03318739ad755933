#pragma once
// multi_gpu.h -- sharding of a benchmark over the GPUs of one box.
//
// One host thread per GPU, every thread drives its own device through the same C ABI on its
// shard of the element range; nothing but ONE double per norm ever crosses NVLink
// (ncclAllReduce of the partial sum of squares).  The partition rule is the one of
// gpu-benchmarking_b200/sharding.py (shard_range): contiguous ranges in units of 32 elements, so
// warp-interleaved groups and 16-byte vectors never straddle GPUs.
// Selected with the environment variable B200FE_NGPUS (default 1: the single-GPU reference flow).
#include <nccl.h>

#include <condition_variable>
#include <mutex>
#include <numeric>
#include <thread>

#include "bench_common.h"

namespace bench
{

#define NCCL_OK(expr)                                                                                        \
    do                                                                                                       \
    {                                                                                                        \
        ncclResult_t r__ = (expr);                                                                           \
        if (r__ != ncclSuccess)                                                                              \
        {                                                                                                    \
            std::fprintf(stderr, "%s:%d: %s failed: %s\n", __FILE__, __LINE__, #expr, ncclGetErrorString(r__)); \
            std::exit(2);                                                                                    \
        }                                                                                                    \
    } while (0)

// [begin, end) of `rank`: mirrors sharding.shard_range (remainder units to the first ranks, tail to the last)
inline void shard_range(size_t total, int rank, int world, size_t multiple, size_t *begin, size_t *end)
{
    const size_t units = total / multiple, tail = total % multiple;
    const size_t base = units / (size_t)world, extra = units % (size_t)world;
    const size_t bu = (size_t)rank * base + std::min<size_t>((size_t)rank, extra);
    const size_t eu = bu + base + ((size_t)rank < extra ? 1 : 0);
    *begin          = bu * multiple;
    *end            = eu * multiple + (rank == world - 1 ? tail : 0);
}

class HostBarrier
{
public:
    explicit HostBarrier(int n) : m_n(n) {}
    void arrive_and_wait()
    {
        std::unique_lock<std::mutex> lock(m_mu);
        const unsigned gen = m_gen;
        if (++m_count == m_n)
        {
            m_count = 0;
            ++m_gen;
            m_cv.notify_all();
        }
        else
            m_cv.wait(lock, [&] { return gen != m_gen; });
    }

private:
    std::mutex m_mu;
    std::condition_variable m_cv;
    int m_n, m_count = 0;
    unsigned m_gen = 0;
};

class MultiGpu
{
public:
    explicit MultiGpu(int n) : m_comms((size_t)n), m_slots((size_t)n, 0.0), m_barrier(n)
    {
        int have = 0;
        CUDA_OK(cudaGetDeviceCount(&have));
        if (n < 1 || n > have)
        {
            std::fprintf(stderr, "B200FE_NGPUS=%d but %d device(s) visible\n", n, have);
            std::exit(2);
        }
        std::vector<int> devs((size_t)n);
        std::iota(devs.begin(), devs.end(), 0);
        NCCL_OK(ncclCommInitAll(m_comms.data(), n, devs.data()));
    }
    ~MultiGpu()
    {
        for (ncclComm_t c : m_comms)
            ncclCommDestroy(c);
    }
    MultiGpu(const MultiGpu &)            = delete;
    MultiGpu &operator=(const MultiGpu &) = delete;
    int size() const { return (int)m_comms.size(); }

    // f(rank) on one thread per GPU, device `rank` current
    template <typename F> void run(F &&f)
    {
        std::vector<std::thread> threads;
        for (int r = 0; r < size(); ++r)
            threads.emplace_back([&, r] {
                CUDA_OK(cudaSetDevice(r));
                f(r);
            });
        for (std::thread &t : threads)
            t.join();
    }
    void barrier() { m_barrier.arrive_and_wait(); }

    // the scalar all-reduce: *d_value (device, one double per rank) -> sum over ranks, returned on every rank
    double allreduce_sum(int rank, double *d_value, cudaStream_t stream = nullptr)
    {
        NCCL_OK(ncclAllReduce(d_value, d_value, 1, ncclDouble, ncclSum, m_comms[(size_t)rank], stream));
        double h = 0.0;
        CUDA_OK(cudaMemcpyAsync(&h, d_value, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        return h;
    }
    // job time of one repetition = the slowest rank (host values, exchanged through shared memory)
    double max_over_ranks(int rank, double v)
    {
        m_slots[(size_t)rank] = v;
        barrier();
        const double m = *std::max_element(m_slots.begin(), m_slots.end());
        barrier();
        return m;
    }

    // min over reps of (max over ranks of host time around fn() + device synchronise), all ranks in step
    template <typename F> double time_min(int rank, unsigned reps, F &&fn)
    {
        Timer t;
        double best = std::numeric_limits<double>::max();
        for (unsigned r = 0; r < reps; ++r)
        {
            barrier();
            t.start();
            fn();
            CUDA_OK(cudaDeviceSynchronize());
            t.stop();
            best = std::min(best, max_over_ranks(rank, t.elapsedSeconds()));
        }
        return best;
    }

private:
    std::vector<ncclComm_t> m_comms;
    std::vector<double> m_slots;
    HostBarrier m_barrier;
};

} // namespace bench

#pragma once
// Host wall-clock timer with the interface of the reference's utils/timer.h
// (start / stop / elapsedNanoseconds / elapsedSeconds): the benchmark logs are
// defined as "min over repetitions of the host time around launch + device
// synchronise".  Written on std::chrono::steady_clock (monotonic); the
// reference mixes system_clock and high_resolution_clock (timer.h:8,42).
#include <chrono>

class Timer
{
public:
    void start()
    {
        m_begin   = clock_type::now();
        m_running = true;
    }

    void stop()
    {
        m_end     = clock_type::now();
        m_running = false;
    }

    double elapsedNanoseconds() const
    {
        const clock_type::time_point until = m_running ? clock_type::now() : m_end;
        return static_cast<double>(std::chrono::duration_cast<std::chrono::nanoseconds>(until - m_begin).count());
    }

    double elapsedSeconds() const
    {
        return elapsedNanoseconds() * 1.0e-9;
    }

private:
    using clock_type = std::chrono::steady_clock;
    clock_type::time_point m_begin{};
    clock_type::time_point m_end{};
    bool m_running = false;
};

// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for opm-core's StopWatch (third party, absent). EulerUpstream_impl.hpp:185-186,214-217.
#ifndef ORACLE_SHIM_STOPWATCH_HPP
#define ORACLE_SHIM_STOPWATCH_HPP
#include <chrono>
namespace Opm { namespace time {
    class StopWatch {
    public:
        StopWatch() : running_(false), elapsed_(0.0) {}
        void start() { t0_ = clock::now(); running_ = true; }
        void stop()  { if (running_) { elapsed_ = secs(clock::now()); running_ = false; } }
        double secsSinceStart() const { return running_ ? secs(clock::now()) : elapsed_; }
        double secsSinceLast() { return secsSinceStart(); }
    private:
        typedef std::chrono::steady_clock clock;
        double secs(clock::time_point t) const { return std::chrono::duration<double>(t - t0_).count(); }
        clock::time_point t0_;
        bool running_;
        double elapsed_;
    };
}}
#endif

// Minimal stand-in for ros_tools/profiling.h: the timers Solver::solve() uses.
#pragma once
#include <chrono>
#include <string>
namespace RosTools {
class Timer {
public:
    explicit Timer(double duration = 0.0) : duration_(duration) {}
    void start() { t0_ = std::chrono::steady_clock::now(); }
    double currentDuration() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); }
    bool hasFinished() const { return currentDuration() >= duration_; }
private:
    double duration_;
    std::chrono::steady_clock::time_point t0_ = std::chrono::steady_clock::now();
};
class Benchmarker {
public:
    explicit Benchmarker(const std::string& = "") {}
    void start() { t0_ = std::chrono::steady_clock::now(); }
    double stop() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); }
private:
    std::chrono::steady_clock::time_point t0_ = std::chrono::steady_clock::now();
};
}  // namespace RosTools
#define PROFILE_SCOPE(x) do { } while (0)
#define PROFILE_FUNCTION() do { } while (0)

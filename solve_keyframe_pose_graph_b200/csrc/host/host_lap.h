#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace pgs {

// PGS_HOST_TIMING=1: wall-clock laps of the host-side phases of a solve (structure analysis, plan, factor allocation) on stderr
struct HostLap {
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  bool on = std::getenv("PGS_HOST_TIMING") != nullptr;
  void lap(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[pgs host] %-34s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

}  // namespace pgs

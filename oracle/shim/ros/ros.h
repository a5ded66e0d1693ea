// ORACLE — TEST INFRASTRUCTURE ONLY.  The handful of roscpp types the reference's NodeDataManager / Worlds /
// PoseGraphSLAM translation units touch (time stamps, the polling rate, a node handle that is only stored).
#pragma once
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <ostream>
#include <string>
#include <thread>
#include <cstdio>
#define ROS_ERROR(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define ROS_WARN(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define ROS_INFO(...) do { } while (0)
#define ROS_ERROR_STREAM(x) do { std::cerr << x << std::endl; } while (0)
#define ROS_WARN_STREAM(x) do { std::cerr << x << std::endl; } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
namespace ros {
struct Duration {
  int32_t sec = 0, nsec = 0;
  Duration() {}
  explicit Duration(double s) { const double f = std::floor(s); sec = (int32_t)f; nsec = (int32_t)std::llround((s - f) * 1e9); if (nsec >= 1000000000) { nsec -= 1000000000; ++sec; } }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  int64_t toNSec() const { return (int64_t)sec * 1000000000LL + (int64_t)nsec; }
  bool operator<(const Duration& o) const { return toNSec() < o.toNSec(); }
  bool operator>(const Duration& o) const { return toNSec() > o.toNSec(); }
  bool operator<=(const Duration& o) const { return toNSec() <= o.toNSec(); }
  bool operator>=(const Duration& o) const { return toNSec() >= o.toNSec(); }
  bool operator==(const Duration& o) const { return toNSec() == o.toNSec(); }
};
struct Time {
  uint32_t sec = 0, nsec = 0;
  Time() {}
  Time(uint32_t s, uint32_t ns) : sec(s), nsec(ns) {}
  explicit Time(double t) { const double f = std::floor(t); sec = (uint32_t)f; nsec = (uint32_t)std::llround((t - f) * 1e9); if (nsec >= 1000000000u) { nsec -= 1000000000u; ++sec; } }
  static Time fromNSec(int64_t ns) { return Time((uint32_t)(ns / 1000000000LL), (uint32_t)(ns % 1000000000LL)); }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  uint64_t toNSec() const { return (uint64_t)sec * 1000000000ull + (uint64_t)nsec; }
  bool isZero() const { return sec == 0 && nsec == 0; }
  Duration operator-(const Time& o) const { Duration d; const int64_t ns = (int64_t)toNSec() - (int64_t)o.toNSec(); int64_t s = ns / 1000000000LL, r = ns % 1000000000LL; if (r < 0) { r += 1000000000LL; --s; } d.sec = (int32_t)s; d.nsec = (int32_t)r; return d; }
  bool operator<(const Time& o) const { return toNSec() < o.toNSec(); }
  bool operator>(const Time& o) const { return toNSec() > o.toNSec(); }
  bool operator<=(const Time& o) const { return toNSec() <= o.toNSec(); }
  bool operator>=(const Time& o) const { return toNSec() >= o.toNSec(); }
  bool operator==(const Time& o) const { return toNSec() == o.toNSec(); }
  bool operator!=(const Time& o) const { return toNSec() != o.toNSec(); }
  static Time now() { return Time(0, 0); }
};
inline std::ostream& operator<<(std::ostream& os, const Time& t) { char b[40]; snprintf(b, sizeof(b), "%u.%09u", t.sec, t.nsec); return os << b; }
inline std::ostream& operator<<(std::ostream& os, const Duration& t) { return os << t.toSec(); }
// Rate::sleep() is where the reference's polling loops yield.  Here it is a GATE, one per thread: the thread reports that it
// arrived (one loop iteration is over) and waits for the test driver to hand it a token, so the driver single-steps each of the
// reference's `while (enabled) { ... rate.sleep(); }` loops deterministically; free_run lets them drain at shutdown.
struct GateState { long tokens = 0, arrivals = 0; };
struct Gate { std::mutex m; std::condition_variable cv; std::map<std::thread::id, GateState> per_thread; bool free_run = false; };
inline Gate& gate() { static Gate g; return g; }
struct Rate {
  double hz;
  explicit Rate(double f) : hz(f) {}
  bool sleep() {
    Gate& g = gate();
    std::unique_lock<std::mutex> lk(g.m);
    GateState& st = g.per_thread[std::this_thread::get_id()];
    ++st.arrivals; g.cv.notify_all();
    g.cv.wait(lk, [&] { return st.tokens > 0 || g.free_run; });
    if (!g.free_run) --st.tokens;
    return true;
  }
};
// publishers swallow what they are given: nothing of what the reference publishes for RViz is part of any comparison
struct Publisher { template <class M> void publish(const M&) const {} int getNumSubscribers() const { return 0; } };
struct NodeHandle {
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  template <class M> Publisher advertise(const std::string&, int, bool = false) const { return Publisher(); }
  template <class T> bool getParam(const std::string&, T&) const { return false; }
};
inline bool ok() { return true; }
}  // namespace ros

// TEST INFRASTRUCTURE: the glog macros gnss_comm uses, as stream sinks (FATAL / failed CHECK abort like glog).
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace glog_shim {
struct Sink {
  bool fatal; std::ostringstream os;
  explicit Sink(bool f) : fatal(f) {}
  ~Sink() { if (fatal) { std::cerr << os.str() << std::endl; std::abort(); } }
  template <typename T> Sink& operator<<(const T& v) { os << v; return *this; }
  Sink& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct Voidify { void operator&(const Sink&) {} };
}  // namespace glog_shim
#define LOG(severity) glog_shim::Sink(glog_shim_is_fatal_##severity)
constexpr bool glog_shim_is_fatal_INFO = false, glog_shim_is_fatal_WARNING = false, glog_shim_is_fatal_ERROR = false,
               glog_shim_is_fatal_FATAL = true;
#define LOG_IF(severity, cond) !(cond) ? (void)0 : glog_shim::Voidify() & LOG(severity)
#define CHECK(cond) (cond) ? (void)0 : glog_shim::Voidify() & glog_shim::Sink(true) << "CHECK failed: " #cond " "
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_NOTNULL(p) (p)
#define DLOG(severity) LOG(severity)
#define VLOG(n) glog_shim::Sink(false)

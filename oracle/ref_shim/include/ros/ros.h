// TEST INFRASTRUCTURE: the few ROS names the reference's hot-path HEADERS mention (no transport, no node).
#pragma once
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

namespace ros {
struct Time {
  double sec_ = 0.0;
  Time() = default;
  explicit Time(double s) : sec_(s) {}
  double toSec() const { return sec_; }
  static Time now() { return Time(); }
};
struct Duration { double sec_ = 0.0; explicit Duration(double s = 0.0) : sec_(s) {} double toSec() const { return sec_; } };
class NodeHandle {
 public:
  NodeHandle() = default;
  explicit NodeHandle(const std::string&) {}
  template <typename T> bool getParam(const std::string&, T&) const { return false; }
  template <typename T> bool param(const std::string&, T& v, const T& d) const { v = d; return false; }
  void shutdown() {}
};
}  // namespace ros
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_WARN_STREAM(x) do { } while (0)
#define ROS_ERROR_STREAM(x) do { std::cerr << x << std::endl; } while (0)
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { } while (0)

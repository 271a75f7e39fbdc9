// TEST INFRASTRUCTURE: feature_tracker/MonoMeas as the plain struct roscpp generates from
// feature_tracker/msg/MonoMeas.msg (uint64 id, float64 image coordinates).
#pragma once
#include <cstdint>
#include <map>      // roscpp message headers bring these in; MapServer.h relies on it
#include <memory>
#include <string>
namespace feature_tracker {
struct MonoMeas {
  uint64_t id = 0;
  double u0 = 0, v0 = 0;
  typedef std::shared_ptr<MonoMeas const> ConstPtr;
  typedef std::shared_ptr<MonoMeas> Ptr;
};
}  // namespace feature_tracker

// TEST INFRASTRUCTURE: gnss_comm/StampedFloat64Array only appears in declarations of gnss_ros.hpp (never compiled here): an empty message
// type with the typedefs roscpp generates is enough for those declarations.
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace gnss_comm {
struct StampedFloat64Array { std_msgs::Header header; typedef std::shared_ptr<StampedFloat64Array const> ConstPtr; typedef std::shared_ptr<StampedFloat64Array> Ptr; };
typedef std::shared_ptr<StampedFloat64Array const> StampedFloat64ArrayConstPtr;
typedef std::shared_ptr<StampedFloat64Array> StampedFloat64ArrayPtr;
}  // namespace gnss_comm

// TEST INFRASTRUCTURE: gnss_comm/GnssTimeMsg only appears in declarations of gnss_ros.hpp (never compiled here): an empty message
// type with the typedefs roscpp generates is enough for those declarations.
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace gnss_comm {
struct GnssTimeMsg { std_msgs::Header header; typedef std::shared_ptr<GnssTimeMsg const> ConstPtr; typedef std::shared_ptr<GnssTimeMsg> Ptr; };
typedef std::shared_ptr<GnssTimeMsg const> GnssTimeMsgConstPtr;
typedef std::shared_ptr<GnssTimeMsg> GnssTimeMsgPtr;
}  // namespace gnss_comm

// TEST INFRASTRUCTURE: gnss_comm/GnssTimePulseInfoMsg only appears in declarations of gnss_ros.hpp (never compiled here): an empty message
// type with the typedefs roscpp generates is enough for those declarations.
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace gnss_comm {
struct GnssTimePulseInfoMsg { std_msgs::Header header; typedef std::shared_ptr<GnssTimePulseInfoMsg const> ConstPtr; typedef std::shared_ptr<GnssTimePulseInfoMsg> Ptr; };
typedef std::shared_ptr<GnssTimePulseInfoMsg const> GnssTimePulseInfoMsgConstPtr;
typedef std::shared_ptr<GnssTimePulseInfoMsg> GnssTimePulseInfoMsgPtr;
}  // namespace gnss_comm

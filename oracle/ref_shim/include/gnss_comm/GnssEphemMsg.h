// TEST INFRASTRUCTURE: gnss_comm/GnssEphemMsg only appears in declarations of gnss_ros.hpp (never compiled here): an empty message
// type with the typedefs roscpp generates is enough for those declarations.
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace gnss_comm {
struct GnssEphemMsg { std_msgs::Header header; typedef std::shared_ptr<GnssEphemMsg const> ConstPtr; typedef std::shared_ptr<GnssEphemMsg> Ptr; };
typedef std::shared_ptr<GnssEphemMsg const> GnssEphemMsgConstPtr;
typedef std::shared_ptr<GnssEphemMsg> GnssEphemMsgPtr;
}  // namespace gnss_comm

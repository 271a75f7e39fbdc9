// TEST INFRASTRUCTURE: gnss_comm/GnssGloEphemMsg only appears in declarations of gnss_ros.hpp (never compiled here): an empty message
// type with the typedefs roscpp generates is enough for those declarations.
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace gnss_comm {
struct GnssGloEphemMsg { std_msgs::Header header; typedef std::shared_ptr<GnssGloEphemMsg const> ConstPtr; typedef std::shared_ptr<GnssGloEphemMsg> Ptr; };
typedef std::shared_ptr<GnssGloEphemMsg const> GnssGloEphemMsgConstPtr;
typedef std::shared_ptr<GnssGloEphemMsg> GnssGloEphemMsgPtr;
}  // namespace gnss_comm

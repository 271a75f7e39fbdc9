// TEST INFRASTRUCTURE: gnss_comm/GnssObsMsg only appears in declarations of gnss_ros.hpp (never compiled here): an empty message
// type with the typedefs roscpp generates is enough for those declarations.
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace gnss_comm {
struct GnssObsMsg { std_msgs::Header header; typedef std::shared_ptr<GnssObsMsg const> ConstPtr; typedef std::shared_ptr<GnssObsMsg> Ptr; };
typedef std::shared_ptr<GnssObsMsg const> GnssObsMsgConstPtr;
typedef std::shared_ptr<GnssObsMsg> GnssObsMsgPtr;
}  // namespace gnss_comm

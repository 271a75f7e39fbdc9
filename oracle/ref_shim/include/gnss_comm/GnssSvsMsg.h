// TEST INFRASTRUCTURE: gnss_comm/GnssSvsMsg only appears in declarations of gnss_ros.hpp (never compiled here): an empty message
// type with the typedefs roscpp generates is enough for those declarations.
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace gnss_comm {
struct GnssSvsMsg { std_msgs::Header header; typedef std::shared_ptr<GnssSvsMsg const> ConstPtr; typedef std::shared_ptr<GnssSvsMsg> Ptr; };
typedef std::shared_ptr<GnssSvsMsg const> GnssSvsMsgConstPtr;
typedef std::shared_ptr<GnssSvsMsg> GnssSvsMsgPtr;
}  // namespace gnss_comm

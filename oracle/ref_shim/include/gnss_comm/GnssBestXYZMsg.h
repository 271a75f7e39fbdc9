// TEST INFRASTRUCTURE: gnss_comm/GnssBestXYZMsg only appears in declarations of gnss_ros.hpp (never compiled here): an empty message
// type with the typedefs roscpp generates is enough for those declarations.
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace gnss_comm {
struct GnssBestXYZMsg { std_msgs::Header header; typedef std::shared_ptr<GnssBestXYZMsg const> ConstPtr; typedef std::shared_ptr<GnssBestXYZMsg> Ptr; };
typedef std::shared_ptr<GnssBestXYZMsg const> GnssBestXYZMsgConstPtr;
typedef std::shared_ptr<GnssBestXYZMsg> GnssBestXYZMsgPtr;
}  // namespace gnss_comm

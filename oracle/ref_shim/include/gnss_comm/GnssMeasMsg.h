// TEST INFRASTRUCTURE: gnss_comm/GnssMeasMsg only appears in declarations of gnss_ros.hpp (never compiled here): an empty message
// type with the typedefs roscpp generates is enough for those declarations.
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace gnss_comm {
struct GnssMeasMsg { std_msgs::Header header; typedef std::shared_ptr<GnssMeasMsg const> ConstPtr; typedef std::shared_ptr<GnssMeasMsg> Ptr; };
typedef std::shared_ptr<GnssMeasMsg const> GnssMeasMsgConstPtr;
typedef std::shared_ptr<GnssMeasMsg> GnssMeasMsgPtr;
}  // namespace gnss_comm

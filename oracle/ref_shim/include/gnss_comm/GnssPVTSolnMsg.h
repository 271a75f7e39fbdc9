// TEST INFRASTRUCTURE: gnss_comm/GnssPVTSolnMsg only appears in declarations of gnss_ros.hpp (never compiled here): an empty message
// type with the typedefs roscpp generates is enough for those declarations.
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace gnss_comm {
struct GnssPVTSolnMsg { std_msgs::Header header; typedef std::shared_ptr<GnssPVTSolnMsg const> ConstPtr; typedef std::shared_ptr<GnssPVTSolnMsg> Ptr; };
typedef std::shared_ptr<GnssPVTSolnMsg const> GnssPVTSolnMsgConstPtr;
typedef std::shared_ptr<GnssPVTSolnMsg> GnssPVTSolnMsgPtr;
}  // namespace gnss_comm

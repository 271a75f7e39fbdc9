// TEST INFRASTRUCTURE: std_msgs/Header as the plain struct roscpp generates (seq, stamp, frame_id).
#pragma once
#include <cstdint>
#include <string>
#include <ros/ros.h>
namespace std_msgs { struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; }; }

// TEST INFRASTRUCTURE: the fields of sensor_msgs/Imu that ImuCtrl reads (ImuPropagator.h:44-49).
#pragma once
#include <memory>
#include <std_msgs/Header.h>
namespace geometry_msgs { struct Vector3 { double x = 0, y = 0, z = 0; }; }
namespace sensor_msgs {
struct Imu {
  std_msgs::Header header;
  geometry_msgs::Vector3 angular_velocity, linear_acceleration;
  typedef std::shared_ptr<Imu const> ConstPtr;
  typedef std::shared_ptr<Imu> Ptr;
};
}  // namespace sensor_msgs

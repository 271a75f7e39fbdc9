// TEST INFRASTRUCTURE: tf::vectorMsgToEigen as used by ImuCtrl (ImuPropagator.h:47-48).
#pragma once
#include <Eigen/Core>
#include <sensor_msgs/Imu.h>
namespace tf {
inline void vectorMsgToEigen(const geometry_msgs::Vector3& m, Eigen::Vector3d& e) { e = Eigen::Vector3d(m.x, m.y, m.z); }
}

// TEST INFRASTRUCTURE: feature_tracker/StereoFrame (header + the list of measurements), msg/StereoFrame.msg.
#pragma once
#include <map>      // roscpp message headers bring these in; MapServer.h relies on it
#include <memory>
#include <string>
#include <vector>
#include <std_msgs/Header.h>
#include <feature_tracker/StereoMeas.h>
namespace feature_tracker {
struct StereoFrame {
  std_msgs::Header header;
  std::vector<StereoMeas> stereo_features;
  typedef std::shared_ptr<StereoFrame const> ConstPtr;
  typedef std::shared_ptr<StereoFrame> Ptr;
};
}  // namespace feature_tracker

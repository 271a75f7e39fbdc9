// TEST INFRASTRUCTURE: feature_tracker/MonoFrame (header + the list of measurements), msg/MonoFrame.msg.
#pragma once
#include <map>      // roscpp message headers bring these in; MapServer.h relies on it
#include <memory>
#include <string>
#include <vector>
#include <std_msgs/Header.h>
#include <feature_tracker/MonoMeas.h>
namespace feature_tracker {
struct MonoFrame {
  std_msgs::Header header;
  std::vector<MonoMeas> mono_features;
  typedef std::shared_ptr<MonoFrame const> ConstPtr;
  typedef std::shared_ptr<MonoFrame> Ptr;
};
}  // namespace feature_tracker

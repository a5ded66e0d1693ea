#pragma once
#include <memory>
#include <string>
#include "../ros/ros.h"
namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; typedef std::shared_ptr<const Header> ConstPtr; };
typedef std::shared_ptr<const Header> HeaderConstPtr;
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
struct Bool { bool data = false; };
}  // namespace std_msgs

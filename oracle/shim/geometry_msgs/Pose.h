// ORACLE — TEST INFRASTRUCTURE ONLY.  Field layout of the geometry_msgs types the reference's sources name.
#pragma once
#include <vector>
#include "../std_msgs/Header.h"
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Point32 { float x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; double covariance[36] = {}; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct PointStamped { std_msgs::Header header; Point point; };
struct Vector3 { double x = 0, y = 0, z = 0; };
}  // namespace geometry_msgs

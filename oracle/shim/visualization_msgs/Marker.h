#pragma once
#include <vector>
#include "../geometry_msgs/Pose.h"
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7, POINTS = 8, TEXT_VIEW_FACING = 9, ADD = 0, DELETE = 2 };
  std_msgs::Header header; std::string ns, text; int id = 0, type = 0, action = 0; geometry_msgs::Pose pose; geometry_msgs::Vector3 scale; std_msgs::ColorRGBA color;
  std::vector<geometry_msgs::Point> points; std::vector<std_msgs::ColorRGBA> colors;
};
struct MarkerArray { std::vector<Marker> markers; };
}  // namespace visualization_msgs

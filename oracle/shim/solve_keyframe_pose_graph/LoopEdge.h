// ORACLE — TEST INFRASTRUCTURE ONLY.  msg/LoopEdge.msg of the reference as a plain struct (the generated header needs catkin).
#pragma once
#include <memory>
#include <string>
#include "../geometry_msgs/Pose.h"
namespace solve_keyframe_pose_graph {
struct LoopEdge { ros::Time timestamp0, timestamp1; geometry_msgs::Pose pose_1T0; float weight = 0; std::string description; typedef std::shared_ptr<const LoopEdge> ConstPtr; };
}

#pragma once
#include "../solve_keyframe_pose_graph/LoopEdge.h"
namespace cerebro {
struct LoopEdge { ros::Time timestamp0, timestamp1; geometry_msgs::Pose pose_1T0; float weight = 0; std::string description; typedef std::shared_ptr<const LoopEdge> ConstPtr; };
}

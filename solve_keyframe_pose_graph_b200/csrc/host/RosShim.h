// ROS message shim: plain structs with the fields of the three messages the reference node
// subscribes to, and the reference's callback names as free functions over pgs::NodeDataManager, so that a ROS node
// forwards its messages unchanged:
//   nav_msgs/Odometry            -> camera_pose_callback            (src/NodeDataManager.cpp:23-103)
//   LoopEdge.msg (msg/LoopEdge.msg: time timestamp0, timestamp1; geometry_msgs/Pose pose_1T0 "pose of 0 as observed
//   from 1"; float32 weight; string description) -> loopclosure_pose_callback (:107-189)
//   std_msgs/Header with frame_id "kidnapped" / "unkidnapped" -> rcvd_kidnap_indicator_callback (:763-792)
// Quaternions arrive as (x,y,z,w) fields and are assembled with Eigen's Quaterniond(w,x,y,z) constructor order, as
// the reference does (:36-41,131-133).  The pose covariance is stored with the keyframe (state file only).
#pragma once
#include <cstdint>
#include <string>

#include "NodeDataManager.h"

namespace pgs {
namespace ros_shim {

struct Time { uint32_t sec = 0, nsec = 0; int64_t toNSec() const { return (int64_t)sec * 1000000000LL + nsec; } };
struct Header { uint32_t seq = 0; Time stamp; std::string frame_id; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; double covariance[36] = {}; };
struct Odometry { Header header; std::string child_frame_id; PoseWithCovariance pose; };   // twist is never read
struct LoopEdge { Time timestamp0, timestamp1; Pose pose_1T0; float weight = 1.0f; std::string description; };

inline Matrix4d pose_to_mat(const Pose& p) {
  const double q[4] = {p.orientation.x, p.orientation.y, p.orientation.z, p.orientation.w};
  const double t[3] = {p.position.x, p.position.y, p.position.z};
  return raw_xyzw_to_mat(q, t);
}

inline void camera_pose_callback(NodeDataManager& m, const Odometry& msg) { m.add_node(msg.header.stamp.toNSec(), pose_to_mat(msg.pose.pose), msg.pose.covariance); }
// false: one of the two timestamps matched no keyframe within 1 ms and the edge was dropped (:181-185)
inline bool loopclosure_pose_callback(NodeDataManager& m, const LoopEdge& msg) {
  return m.add_loop_edge(msg.timestamp0.toNSec(), msg.timestamp1.toNSec(), pose_to_mat(msg.pose_1T0), (double)msg.weight, msg.description);
}
// false: frame_id is neither "kidnapped" nor "unkidnapped" (the reference exits, :789-791) or the state transition is invalid
inline bool rcvd_kidnap_indicator_callback(NodeDataManager& m, const Header& h) {
  if (h.frame_id == "kidnapped") return m.rcvd_kidnap_indicator(h.stamp.toNSec(), true);
  if (h.frame_id == "unkidnapped") return m.rcvd_kidnap_indicator(h.stamp.toNSec(), false);
  return false;
}

}  // namespace ros_shim
}  // namespace pgs

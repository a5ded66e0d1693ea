// ROS-free drop-in for the pose-assembly part of the reference's Composer (reference src/Composer.h:52-84,
// src/Composer.cpp:10-292): same constructor shape (manager, slam), the same thread entry point with its
// enable/disable switches, the same outputs (global_jmb: world -> poses, global_lmb: pose per keyframe,
// latest world id) and get_last_known_camerapose().  The per-keyframe arithmetic of the loop body runs as one
// CUDA gather through include/pgs_compose.h; the publishers (bf_traj / cam_visual / path / loopedge threads,
// Composer.cpp:294-1100) are ROS visualisation and stay out of scope.
#pragma once
#include <atomic>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/pgs_compose.h"
#include "NodeDataManager.h"
#include "PoseGraphSLAM.h"

namespace pgs {

class Composer {
 public:
  Composer(const NodeDataManager* manager, const PoseGraphSLAM* slam, int device = 0);
  ~Composer();

  // 30 Hz loop of the reference (Composer.cpp:10-263); runs until pose_assember_disable().
  void pose_assember_thread(int looprate = 30);
  void pose_assember_enable() { b_pose_assember = true; }
  void pose_assember_disable() { b_pose_assember = false; }
  // One pass of the loop body (Composer.cpp:24-255).  false + last_error() on failure; true (and no change) when
  // the manager has no keyframes yet (:26-30).
  bool pose_assember_once();
  const std::string& last_error() const { return error_; }

  // w_T_lastcam and its timestamp (ns); returns the index of that keyframe, or -1 before the first pass (Composer.cpp:264-276)
  int get_last_known_camerapose(Matrix4d& w_T_lastcam, int64_t& stamp_of_it) const;

  // copies of the published state (the reference's publisher threads read these under the same mutex)
  std::map<int, std::vector<Matrix4d>> get_global_jmb() const;
  std::vector<Matrix4d> get_global_lmb() const;
  int get_global_latest_pose_worldid() const;
  double last_kernel_ms() const { return ms_kernel_; }
  double last_total_ms() const { return ms_total_; }

 private:
  mutable std::mutex mx;
  const NodeDataManager* manager;
  const PoseGraphSLAM* slam;
  pgs_compose_handle handle_ = nullptr;
  int device_ = 0;
  std::atomic<bool> b_pose_assember;
  std::map<int, std::vector<Matrix4d>> global_jmb;   // key: worldID, value: vector of poses
  std::vector<Matrix4d> global_lmb;                  // corrected poses, same index as the node
  int global_latest_pose_worldid = -1;
  std::string error_;
  double ms_kernel_ = 0, ms_total_ = 0;
  // staging (reused between passes)
  std::vector<double> mgr_T_, slam_q_, slam_t_, ws_T_w_, out_T_;
  std::vector<int32_t> world_id_, world_end_, world_setid_;
  std::vector<uint8_t> ws_exists_;
};

}  // namespace pgs

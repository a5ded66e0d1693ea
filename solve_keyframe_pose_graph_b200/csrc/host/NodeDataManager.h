// ROS-free node / loop-edge store with the getter surface the solver front-end consumes
// (reference src/NodeDataManager.h:95-111 node/edge getters, :152-187 kidnap/world queries).
// ros::Time becomes int64 nanoseconds; the ROS callbacks become plain methods:
//   camera_pose_callback          (NodeDataManager.cpp:23-103)  -> add_node()
//   loopclosure_pose_callback     (NodeDataManager.cpp:107-189) -> add_loop_edge() (timestamp lookup, 1 ms tolerance)
//   rcvd_kidnap_indicator_callback(NodeDataManager.cpp:763-792) -> rcvd_kidnap_indicator()
// Covariances are stored and round-trip through the state file but are never read by the solver; extrinsics are out of scope.
#pragma once
#include <array>
#include <atomic>
#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "Worlds.h"
#include "pose_math.h"

namespace pgs {

class NodeDataManager {
 public:
  NodeDataManager();
  ~NodeDataManager();

  // ---- ingest
  // cov36: row-major 6x6 pose covariance of the Odometry message (kept for the state file only; the solver never reads it)
  void add_node(int64_t stamp_ns, const Matrix4d& w_T_cam, const double* cov36 = nullptr);
  // Returns false (edge dropped) when either timestamp matches no node, as the reference does (:181-185).
  bool add_loop_edge(int64_t stamp_a_ns, int64_t stamp_b_ns, const Matrix4d& b_T_a, double weight, const std::string& description = "");
  bool add_loop_edge_by_index(int a, int b, const Matrix4d& b_T_a, double weight, const std::string& description = "");
  bool rcvd_kidnap_indicator(int64_t stamp_ns, bool kidnapped);

  // ---- on-disk state (NodeDataManager.cpp:503-754); defined in GraphIO.cpp
  bool saveAsJSON(const std::string& base_path) const;                                             // log_posegraph.json
  bool loadFromJSON(const std::string& base_path, const std::vector<bool>& edge_mask = {});      // + kidnap signals replayed

  // ---- node getters
  int getNodeLen() const;
  bool getNodePose(int i, Matrix4d& w_T_cam) const;
  const Matrix4d& getNodePose(int i) const;
  bool nodePoseExists(int i) const;
  bool getNodeCov(int i, double* cov36) const;               // reference NodeDataManager.cpp:363-381
  int64_t getNodeTimestamp(int i) const;
  // Keyframes [from, to) in one locked pass: odometry pose and which_world_is_this(stamp) of each (what the trigger otherwise asks
  // for keyframe by keyframe, a lock per call)
  void snapshot_nodes(int from, int to, std::vector<Matrix4d>& poses, std::vector<int>& world_of) const;
  // ---- edge getters
  int getEdgeLen() const;
  const Matrix4d& getEdgePose(int i) const;                 // b_T_a
  const std::pair<int, int>& getEdgeIdxInfo(int i) const;   // (a, b)
  double getEdgeWeight(int i) const;
  const std::string getEdgeDescriptionString(int i) const;

  // ---- restoring a saved session (the two manager steps of Composer::loadStateFromDisk, reference src/Composer.cpp:1148-1163;
  // the Worlds object must have been restored first).  load_kidnap_data: NodeDataManager::load_kidnap_data_from_json
  // (NodeDataManager.cpp:912-950) — the stamp lists replace the current ones, the kidnap status follows from their lengths;
  // false when they cannot belong together (the reference exits).  load_solved_node: one entry of "SolvedPoseGraph" as
  // NodeDataManager::load_solved_posegraph_data_from_json treats it (:998-1090): the saved pose is in the frame of its world's
  // set root and is moved back into the world's own frame, the stamp must fall into the saved world and set; no world is started.
  bool load_kidnap_data(const std::vector<int64_t>& starts_ns, const std::vector<int64_t>& ends_ns);
  // NodeDataManager::mark_as_kidnapped_and_signal_end_of_world (NodeDataManager.cpp:838-844): what Composer::saveStateToDisk does before it
  // writes when the session is not kidnapped — the current world ends at the stamp of the last keyframe.  false: no keyframes / already kidnapped.
  bool mark_as_kidnapped_and_signal_end_of_world();
  bool load_solved_node(int64_t stamp_ns, const Matrix4d& ws_T_c, int world_id, int set_id_of_world, std::string* err = nullptr);

  // ---- kidnap / world queries
  bool curr_kidnap_status() const { return current_kidnap_status; }
  int n_kidnaps() const;
  int64_t stamp_of_kidnap_i_started(int i) const;
  int64_t stamp_of_kidnap_i_ended(int i) const;
  int nodeidx_of_world_i_started(int i) const;
  int nodeidx_of_world_i_ended(int i) const;
  int n_worlds() const;
  // >= 0: world id; negative: kidnap dead zone -(k+1)   (NodeDataManager.cpp:1127-1198)
  int which_world_is_this(int64_t stamp_ns) const;
  Worlds* getWorldsPtr() { return worlds_handle_raw_ptr; }
  const Worlds* getWorldsConstPtr() const { return worlds_handle_raw_ptr; }

 private:
  int find_indexof_node(int64_t stamp_ns) const;   // caller holds node_mutex
  int which_world_nolock(int64_t t) const;

  mutable std::mutex node_mutex;
  // deques: the getters hand out references (as the reference's do) while the ingest thread appends; a deque's
  // push_back never moves existing elements, a vector's reallocation would leave those references dangling
  std::deque<Matrix4d> node_pose;
  std::deque<int64_t> node_timestamps;
  std::deque<std::array<double, 36>> node_pose_covariance;

  mutable std::mutex edge_mutex;
  std::deque<std::pair<int, int>> loopclosure_edges;
  std::deque<double> loopclosure_edges_goodness;
  std::deque<Matrix4d> loopclosure_p_T_c;
  std::deque<std::string> loopclosure_description;

  mutable std::mutex mutex_kidnap;
  std::vector<int64_t> kidnap_starts, kidnap_ends;
  std::atomic<bool> current_kidnap_status;
  bool first_keyframe_received = false;   // guarded by node_mutex
  Worlds* worlds_handle_raw_ptr = nullptr;
};

}  // namespace pgs

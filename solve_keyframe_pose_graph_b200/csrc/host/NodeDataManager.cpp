#include "NodeDataManager.h"

#include <algorithm>

namespace pgs {

NodeDataManager::NodeDataManager() : current_kidnap_status(false) { worlds_handle_raw_ptr = new Worlds(); }
NodeDataManager::~NodeDataManager() { delete worlds_handle_raw_ptr; }

void NodeDataManager::add_node(int64_t stamp_ns, const Matrix4d& w_T_cam, const double* cov36) {
  std::lock_guard<std::mutex> lk(node_mutex);
  node_timestamps.push_back(stamp_ns);
  node_pose.push_back(w_T_cam);
  std::array<double, 36> cov{};                                                 // NodeDataManager.cpp:55-63
  if (cov36) for (int k = 0; k < 36; ++k) cov[k] = cov36[k];
  node_pose_covariance.push_back(cov);
  // the first keyframe this manager RECEIVES (NodeDataManager.cpp:74-92; the reference keeps the flag in a function-local static):
  // on a fresh start world 0 begins; after a restore from disk — the file was saved with the world ended, the session is kidnapped —
  // the kidnap ends and a new world begins at this keyframe (mark_as_unkidnapped_and_signal_start_of_world)
  if (!first_keyframe_received) {
    first_keyframe_received = true;
    if (node_pose.size() == 1) worlds_handle_raw_ptr->world_starts(stamp_ns);
    else if (current_kidnap_status) {
      { std::lock_guard<std::mutex> lk2(mutex_kidnap); current_kidnap_status = false; kidnap_ends.push_back(stamp_ns); }
      worlds_handle_raw_ptr->world_starts(stamp_ns);
    }
  }
}

// First node whose stamp is strictly within 1 ms of `stamp` (the reference scans linearly and returns
// the first hit, NodeDataManager.cpp:274-299); stamps are increasing, so a binary search finds the same index.
int NodeDataManager::find_indexof_node(int64_t stamp_ns) const {
  const int64_t tol = 1000000;
  auto it = std::upper_bound(node_timestamps.begin(), node_timestamps.end(), stamp_ns - tol);
  if (it == node_timestamps.end()) return -1;
  return (*it < stamp_ns + tol) ? (int)(it - node_timestamps.begin()) : -1;
}

bool NodeDataManager::add_loop_edge(int64_t stamp_a_ns, int64_t stamp_b_ns, const Matrix4d& b_T_a, double weight, const std::string& description) {
  int ia, ib;
  { std::lock_guard<std::mutex> lk(node_mutex); ia = find_indexof_node(stamp_a_ns); ib = find_indexof_node(stamp_b_ns); }
  if (ia < 0 || ib < 0) return false;
  return add_loop_edge_by_index(ia, ib, b_T_a, weight, description);
}

bool NodeDataManager::add_loop_edge_by_index(int a, int b, const Matrix4d& b_T_a, double weight, const std::string& description) {
  if (a < 0 || b < 0 || a >= getNodeLen() || b >= getNodeLen()) return false;
  std::lock_guard<std::mutex> lk(edge_mutex);
  loopclosure_edges.push_back({a, b});
  loopclosure_edges_goodness.push_back(weight);
  loopclosure_p_T_c.push_back(b_T_a);
  loopclosure_description.push_back(description);
  return true;
}

bool NodeDataManager::rcvd_kidnap_indicator(int64_t stamp_ns, bool kidnapped) {
  if (kidnapped) {
    if (current_kidnap_status) return false;
    { std::lock_guard<std::mutex> lk(mutex_kidnap); current_kidnap_status = true; kidnap_starts.push_back(stamp_ns); }
    worlds_handle_raw_ptr->world_ends(stamp_ns);
  } else {
    if (!current_kidnap_status) return false;
    { std::lock_guard<std::mutex> lk(mutex_kidnap); current_kidnap_status = false; kidnap_ends.push_back(stamp_ns); }
    worlds_handle_raw_ptr->world_starts(stamp_ns);
  }
  return true;
}

int NodeDataManager::getNodeLen() const { std::lock_guard<std::mutex> lk(node_mutex); return (int)node_pose.size(); }
bool NodeDataManager::getNodePose(int i, Matrix4d& w_T_cam) const {
  std::lock_guard<std::mutex> lk(node_mutex);
  if (i < 0 || i >= (int)node_pose.size()) return false;
  w_T_cam = node_pose[i]; return true;
}
const Matrix4d& NodeDataManager::getNodePose(int i) const { std::lock_guard<std::mutex> lk(node_mutex); return node_pose[i]; }
bool NodeDataManager::mark_as_kidnapped_and_signal_end_of_world() {
  int64_t last;
  { std::lock_guard<std::mutex> lk(node_mutex); if (node_timestamps.empty()) return false; last = node_timestamps.back(); }
  return rcvd_kidnap_indicator(last, true);
}

bool NodeDataManager::load_kidnap_data(const std::vector<int64_t>& starts_ns, const std::vector<int64_t>& ends_ns) {
  if (!(starts_ns.size() == ends_ns.size() || starts_ns.size() == ends_ns.size() + 1)) return false;      // the reference exit(1)s (:944-948)
  std::lock_guard<std::mutex> lk(mutex_kidnap);
  kidnap_starts = starts_ns; kidnap_ends = ends_ns;
  current_kidnap_status = starts_ns.size() != ends_ns.size();
  return true;
}

bool NodeDataManager::load_solved_node(int64_t stamp_ns, const Matrix4d& ws_T_c, int world_id, int set_id, std::string* err) {
  auto fail = [&](const std::string& m) { if (err) *err = m; return false; };
  Matrix4d w_T_c = ws_T_c;
  if (world_id >= 0 && world_id != set_id) {                                                            // :1038-1050
    bool ok = false;
    const Matrix4d w_T_ws = worlds_handle_raw_ptr->is_exist(world_id, set_id) ? worlds_handle_raw_ptr->getPoseBetweenWorlds(world_id, set_id, &ok) : Matrix4d::Identity();
    if (!ok) return fail("no relative pose between world " + std::to_string(world_id) + " and its saved set root " + std::to_string(set_id));
    w_T_c = w_T_ws * ws_T_c;
  }
  if (world_id >= 0 && set_id >= 0) {                                                                   // :1061-1071
    const int w = which_world_is_this(stamp_ns);
    if (w != world_id) return fail("a saved keyframe falls into world " + std::to_string(w) + ", the file says " + std::to_string(world_id));
    if (worlds_handle_raw_ptr->find_setID_of_world_i(w) != set_id) return fail("a saved keyframe's world is in another set than the file says");
  }
  std::lock_guard<std::mutex> lk(node_mutex);
  node_timestamps.push_back(stamp_ns); node_pose.push_back(w_T_c); node_pose_covariance.push_back(std::array<double, 36>{});   // :1078-1081, no world_starts()
  return true;
}

bool NodeDataManager::getNodeCov(int i, double* cov36) const {
  std::lock_guard<std::mutex> lk(node_mutex);
  if (i < 0 || i >= (int)node_pose_covariance.size() || !cov36) return false;
  for (int k = 0; k < 36; ++k) cov36[k] = node_pose_covariance[i][k];
  return true;
}
bool NodeDataManager::nodePoseExists(int i) const { std::lock_guard<std::mutex> lk(node_mutex); return i >= 0 && i < (int)node_pose.size(); }
int64_t NodeDataManager::getNodeTimestamp(int i) const {
  std::lock_guard<std::mutex> lk(node_mutex);
  if (i < 0 || i >= (int)node_timestamps.size()) return -1;
  return node_timestamps[i];
}
void NodeDataManager::snapshot_nodes(int from, int to, std::vector<Matrix4d>& poses, std::vector<int>& world_of) const {
  std::vector<int64_t> stamps;
  {
    std::lock_guard<std::mutex> lk(node_mutex);
    from = std::max(from, 0); to = std::min(to, (int)node_pose.size());
    const int n = std::max(0, to - from);
    poses.assign(node_pose.begin() + from, node_pose.begin() + from + n);
    stamps.assign(node_timestamps.begin() + from, node_timestamps.begin() + from + n);
  }
  std::lock_guard<std::mutex> lk(mutex_kidnap);
  world_of.resize(stamps.size());
  for (size_t i = 0; i < stamps.size(); ++i) world_of[i] = which_world_nolock(stamps[i]);
}
int NodeDataManager::getEdgeLen() const { std::lock_guard<std::mutex> lk(edge_mutex); return (int)loopclosure_edges.size(); }
const Matrix4d& NodeDataManager::getEdgePose(int i) const { std::lock_guard<std::mutex> lk(edge_mutex); return loopclosure_p_T_c[i]; }
const std::pair<int, int>& NodeDataManager::getEdgeIdxInfo(int i) const { std::lock_guard<std::mutex> lk(edge_mutex); return loopclosure_edges[i]; }
double NodeDataManager::getEdgeWeight(int i) const {
  std::lock_guard<std::mutex> lk(edge_mutex);
  return (i >= 0 && i < (int)loopclosure_edges_goodness.size()) ? loopclosure_edges_goodness[i] : -1.0;   // NodeDataManager.cpp:458-466
}
const std::string NodeDataManager::getEdgeDescriptionString(int i) const {
  std::lock_guard<std::mutex> lk(edge_mutex);
  return (i >= 0 && i < (int)loopclosure_description.size()) ? loopclosure_description[i] : std::string("NA");
}

int NodeDataManager::n_kidnaps() const { std::lock_guard<std::mutex> lk(mutex_kidnap); return (int)kidnap_ends.size(); }
int64_t NodeDataManager::stamp_of_kidnap_i_started(int i) const {
  std::lock_guard<std::mutex> lk(mutex_kidnap);
  return (i >= 0 && i < (int)kidnap_starts.size()) ? kidnap_starts[i] : 0;
}
int64_t NodeDataManager::stamp_of_kidnap_i_ended(int i) const {
  std::lock_guard<std::mutex> lk(mutex_kidnap);
  return (i >= 0 && i < (int)kidnap_ends.size()) ? kidnap_ends[i] : 0;
}
int NodeDataManager::n_worlds() const { std::lock_guard<std::mutex> lk(mutex_kidnap); return (int)kidnap_ends.size() + 1; }

int NodeDataManager::which_world_is_this(int64_t t) const { std::lock_guard<std::mutex> lk(mutex_kidnap); return which_world_nolock(t); }

// Restates NodeDataManager.cpp:1127-1198 branch for branch, including its boundary conventions
// (the single-kidnap branch uses >= / <= where the general branch uses > / <=).
int NodeDataManager::which_world_nolock(int64_t t) const {
  const size_t ns = kidnap_starts.size(), ne = kidnap_ends.size();
  if (ns == 0) return 0;
  if (ns == 1) {
    if (t < kidnap_starts[0]) return 0;
    if (ne == 0) return -1;
    return (t >= kidnap_starts[0] && t <= kidnap_ends[0]) ? -1 : 1;
  }
  int64_t prev = 0;   // ros::Time()
  if (ns == ne) {
    for (size_t i = 0; i < ns; ++i) {
      if (t > prev && t <= kidnap_starts[i]) return (int)i;
      if (t > kidnap_starts[i] && t <= kidnap_ends[i]) return -((int)i + 1);
      prev = kidnap_ends[i];
    }
    return (int)ne;
  }
  // currently kidnapped: one more start than ends
  for (size_t i = 0; i + 1 < ns; ++i) {
    if (t > prev && t <= kidnap_starts[i]) return (int)i;
    if (t > kidnap_starts[i] && t <= kidnap_ends[i]) return -((int)i + 1);
    prev = kidnap_ends[i];
  }
  const int i = (int)ns - 1;
  if (t > kidnap_ends[i - 1] && t <= kidnap_starts[i]) return i;
  // the reference only answers for t > kidnap_starts[i] and otherwise falls off the end (UB); every
  // remaining stamp is treated as inside the open dead zone.
  return -(i + 1);
}

int NodeDataManager::nodeidx_of_world_i_started(int i) const {   // NodeDataManager.cpp:1213-1250
  if (i < 0) return -3;
  if (i == 0) return 0;
  int n;
  { std::lock_guard<std::mutex> lk(mutex_kidnap); n = (int)kidnap_ends.size(); }
  if (i - 1 < n) {
    std::lock_guard<std::mutex> lk(node_mutex);
    std::lock_guard<std::mutex> lk2(mutex_kidnap);
    for (size_t r = 0; r < node_timestamps.size(); ++r) if (which_world_nolock(node_timestamps[r]) == i) return (int)r;
  }
  return -4;
}

int NodeDataManager::nodeidx_of_world_i_ended(int i) const {     // NodeDataManager.cpp:1256-1292
  int n_ends, n_starts; int64_t start_i = 0;
  { std::lock_guard<std::mutex> lk(mutex_kidnap); n_ends = (int)kidnap_ends.size(); n_starts = (int)kidnap_starts.size(); if (i >= 0 && i < n_starts) start_i = kidnap_starts[i]; }
  if (i < 0 || i > n_ends) return -1;
  std::lock_guard<std::mutex> lk(node_mutex);
  if (i < n_starts) return find_indexof_node(start_i);   // -1 when no keyframe lies within 1 ms of the kidnap stamp
  return (int)node_timestamps.size() - 1;
}

}  // namespace pgs

#include "Composer.h"

#include <chrono>
#include <cstring>
#include <thread>

namespace pgs {

Composer::Composer(const NodeDataManager* _manager, const PoseGraphSLAM* _slam, int device) : manager(_manager), slam(_slam), device_(device) {
  b_pose_assember = false;
}

Composer::~Composer() { if (handle_) pgs_compose_destroy(handle_); }

void Composer::pose_assember_thread(int looprate) {
  if (looprate <= 0 || looprate >= 50) looprate = 30;   // the reference asserts 0 < looprate < 50 (Composer.cpp:13)
  const auto period = std::chrono::microseconds(1000000 / looprate);
  while (b_pose_assember) {
    const auto t0 = std::chrono::steady_clock::now();
    pose_assember_once();
    std::this_thread::sleep_until(t0 + period);
  }
}

bool Composer::pose_assember_once() {
  error_.clear();
  const int n = manager->getNodeLen();
  if (n == 0) return true;                                                          // Composer.cpp:26-30
  if (!handle_) {
    const int rc = pgs_compose_create(device_, &handle_);
    if (rc != PGS_OK) { error_ = "pgs_compose_create failed: no usable CUDA device (there is no CPU fallback)"; handle_ = nullptr; return false; }
  }
  // ---- snapshot of what the loop body reads (Composer.cpp:35-58): solver first, so that solved_until never
  // points past the poses we copy
  int solved_until = slam->solvedUntil();
  slam->getAllNodeRaw(slam_q_, slam_t_);
  const int n_slam = (int)(slam_t_.size() / 3);
  if (solved_until < 0) solved_until = 0;
  if (solved_until >= n) solved_until = n - 1;
  const int solved_until_world = manager->which_world_is_this(manager->getNodeTimestamp(solved_until));
  mgr_T_.resize(16 * (size_t)n); world_id_.resize(n);
  for (int i = 0; i < n; ++i) {
    std::memcpy(&mgr_T_[16 * (size_t)i], manager->getNodePose(i).m, 128);
    world_id_[i] = manager->which_world_is_this(manager->getNodeTimestamp(i));
  }
  const int nw = manager->n_worlds();
  const Worlds* W = manager->getWorldsConstPtr();
  world_end_.resize(nw); world_setid_.resize(nw); ws_exists_.resize(nw); ws_T_w_.assign(16 * (size_t)nw, 0.0);
  for (int w = 0; w < nw; ++w) {
    world_end_[w] = manager->nodeidx_of_world_i_ended(w);
    const int setid = W->find_setID_of_world_i(w);
    world_setid_[w] = setid;
    bool ok = false;
    Matrix4d T = Matrix4d::Identity();
    if (setid != w && W->is_exist(setid, w)) T = W->getPoseBetweenWorlds(setid, w, &ok);   // Composer.cpp:177-183
    ws_exists_[w] = ok ? 1 : 0;
    std::memcpy(&ws_T_w_[16 * (size_t)w], T.m, 128);
  }
  out_T_.resize(16 * (size_t)n);
  pgs_compose_input in;
  in.n_nodes = n; in.mgr_T = mgr_T_.data(); in.world_id = world_id_.data();
  in.n_slam = std::min(n_slam, n); in.slam_q = slam_q_.data(); in.slam_t = slam_t_.data();
  in.solved_until = solved_until; in.solved_until_world = solved_until_world;
  in.n_worlds = nw; in.world_end = world_end_.data(); in.world_setid = world_setid_.data(); in.ws_exists = ws_exists_.data(); in.ws_T_w = ws_T_w_.data();
  const int rc = pgs_compose_run(handle_, &in, out_T_.data());
  if (rc != PGS_OK) { error_ = pgs_compose_last_error(handle_); return false; }
  pgs_compose_last_timing(handle_, &ms_kernel_, &ms_total_);
  // ---- publish (Composer.cpp:211-255): jmb groups the poses by world in keyframe order, lmb is the flat list
  std::map<int, std::vector<Matrix4d>> jmb;
  std::vector<Matrix4d> lmb(n);
  for (int i = 0; i < n; ++i) {
    std::memcpy(lmb[i].m, &out_T_[16 * (size_t)i], 128);
    jmb[world_id_[i]].push_back(lmb[i]);
  }
  {
    std::lock_guard<std::mutex> lk(mx);
    global_jmb.swap(jmb);
    global_lmb.swap(lmb);
    global_latest_pose_worldid = world_id_[n - 1];
  }
  return true;
}

int Composer::get_last_known_camerapose(Matrix4d& w_T_lastcam, int64_t& stamp_of_it) const {
  std::lock_guard<std::mutex> lk(mx);
  const int sz = (int)global_lmb.size();
  if (sz == 0) return -1;                                     // Composer.cpp:268-269
  w_T_lastcam = global_lmb[sz - 1];
  stamp_of_it = manager->getNodeTimestamp(sz - 1);
  return sz - 1;                                              // :274-275
}

std::map<int, std::vector<Matrix4d>> Composer::get_global_jmb() const { std::lock_guard<std::mutex> lk(mx); return global_jmb; }
std::vector<Matrix4d> Composer::get_global_lmb() const { std::lock_guard<std::mutex> lk(mx); return global_lmb; }
int Composer::get_global_latest_pose_worldid() const { std::lock_guard<std::mutex> lk(mx); return global_latest_pose_worldid; }

}  // namespace pgs

// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// THE REFERENCE'S OWN FRONT END, single-stepped.  src/NodeDataManager.cpp, src/Worlds.cpp, src/PoseGraphSLAM.cpp, src/Composer.cpp
// (with src/VizPoseGraph.cpp, whose publishers are inert), src/utils/PoseManipUtils.cpp, src/utils/RawFileIO.cpp and
// src/utils/RosMarkerUtils.cpp are compiled unmodified, from where they lie under /root/reference, over oracle/shim/
// (stand-ins for the Eigen / Ceres / roscpp / message / OpenCV names they use) into oracle/_ref/libref_frontend.so.
// The ROS callbacks are fed messages built from plain arrays; PoseGraphSLAM::reinit_ceres_problem_onnewloopedge_
// optimize6DOF() runs on its own thread exactly as in keyframe_pose_graph_slam_node.cpp:475-477, with ros::Rate::sleep
// turned into a gate so that one call of refslam_wakeup() is one wake-up of the reference's loop.  The shim's
// ceres::Problem records what the reference builds; ceres::Solve minimises nothing — it snapshots the problem and the
// optimisation variables (= the initial guesses the reference just wrote).  tests/test_reference_frontend.py compares
// that, wake-up by wake-up, with what the product's facade and the oracle's Python front-end build from the same input:
// which residual blocks exist, between which keyframes, with which observation and weight, which keyframes get
// regularisers, every initial guess, solvedUntil, and the world / set bookkeeping.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <typeinfo>
#include <vector>

#include "nlohmann/json.hpp"    // (and every standard header above) before the access widening below
#include "Eigen/Dense"
#include "ceres/ceres.h"
#include "ros/ros.h"
#include "opencv2/core/core.hpp"

#define private public          // the functors keep their observation and weight private; this translation unit only reads them
#define protected public
#include "PoseGraphSLAM.h"      // -I /root/reference/src
#include "Composer.h"
#undef private
#undef protected

namespace {

struct Block { int type, c1, c2, sw; double obs[16]; double weight; };   // type 0 SixDOFError, 1 ...WithSwitchingConstraints, 2 NodePoseRegularization
struct Snapshot { std::vector<Block> blocks; std::vector<double> q, t, s; std::vector<int> constant; int unknown_blocks = 0; };

struct Ref {
  ros::NodeHandle nh;
  NodeDataManager* manager = nullptr;
  PoseGraphSLAM* slam = nullptr;
  Composer* composer = nullptr;
  std::thread th, th_composer;
  bool started = false, composer_started = false;
  std::vector<Snapshot> snaps;
  long arrivals_seen = 0, composer_arrivals_seen = 0;
  void (*solve_callback)() = nullptr;   // called inside the stand-in ceres::Solve, after the snapshot: the test's solver
  double perturb = 0.0;          // > 0: the stand-in "solve" moves every variable of the problem by a known, index-dependent amount
};
Ref* g_ref = nullptr;

void copy16(const Matrix4d& M, double* o) { for (int i = 0; i < 16; ++i) o[i] = M.a[i]; }

// called by the shim's ceres::Solve on the solver thread
void on_solve(const ceres::Solver::Options&, ceres::Problem* P, ceres::Solver::Summary* sum) {
  Ref* R = g_ref;
  Snapshot S;
  const int n = R->slam->n_opt_variables(), ns = R->slam->n_opt_switch();
  std::map<const double*, int> qi, ti, si;
  for (int i = 0; i < n; ++i) { qi[R->slam->get_raw_ptr_to_opt_variable_q(i)] = i; ti[R->slam->get_raw_ptr_to_opt_variable_t(i)] = i; }
  for (int e = 0; e < ns; ++e) si[R->slam->get_raw_ptr_to_opt_switch(e)] = e;
  for (ceres::ResidualBlock* b : P->blocks) {
    if (b->removed) continue;
    Block B; std::memset(&B, 0, sizeof(B)); B.c1 = B.c2 = B.sw = -1;
    const std::string name = b->cost->functor_name();
    if (name == typeid(SixDOFError).name() && b->params.size() == 4) {
      const SixDOFError* f = (const SixDOFError*)b->cost->functor();
      B.type = 0; B.c1 = qi.at(b->params[0]); B.c2 = qi.at(b->params[2]); copy16(f->observed__c1_T_c2, B.obs); B.weight = f->weight;
      if (ti.at(b->params[1]) != B.c1 || ti.at(b->params[3]) != B.c2) ++S.unknown_blocks;
    } else if (name == typeid(SixDOFErrorWithSwitchingConstraints).name() && b->params.size() == 5) {
      const SixDOFErrorWithSwitchingConstraints* f = (const SixDOFErrorWithSwitchingConstraints*)b->cost->functor();
      B.type = 1; B.c1 = qi.at(b->params[0]); B.c2 = qi.at(b->params[2]); B.sw = si.at(b->params[4]); copy16(f->observed__c1_T_c2, B.obs); B.weight = f->weight;
      if (ti.at(b->params[1]) != B.c1 || ti.at(b->params[3]) != B.c2) ++S.unknown_blocks;
    } else if (name == typeid(NodePoseRegularization).name() && b->params.size() == 2) {
      const NodePoseRegularization* f = (const NodePoseRegularization*)b->cost->functor();
      B.type = 2; B.c1 = qi.at(b->params[0]); copy16(f->nodepose, B.obs); B.weight = f->weight;
      if (ti.at(b->params[1]) != B.c1) ++S.unknown_blocks;
    } else { ++S.unknown_blocks; continue; }
    S.blocks.push_back(B);
  }
  S.q.resize(4 * (size_t)n); S.t.resize(3 * (size_t)n); S.s.resize(ns); S.constant.assign(n, 0);
  for (int i = 0; i < n; ++i) {
    const double* q = R->slam->get_raw_ptr_to_opt_variable_q(i); const double* t = R->slam->get_raw_ptr_to_opt_variable_t(i);
    for (int k = 0; k < 4; ++k) S.q[4 * i + k] = q[k];
    for (int k = 0; k < 3; ++k) S.t[3 * i + k] = t[k];
    auto it = P->param_const.find(const_cast<double*>(q)); S.constant[i] = it != P->param_const.end() && it->second;
  }
  for (int e = 0; e < ns; ++e) S.s[e] = *R->slam->get_raw_ptr_to_opt_switch(e);
  sum->termination_type = ceres::NO_CONVERGENCE;
  R->snaps.push_back(S);
  if (R->solve_callback) R->solve_callback();   // may call refslam_write_vars(): that is "the solve"
  if (R->perturb > 0) {
    // A stand-in for what a solve does to the state the NEXT wake-up starts from: every pose that appears in a residual
    // block moves by Plus(q, a*eps_i), t += a*d_i, every switch of a block changes, all as closed forms of the index and
    // the wake-up number so that tests/test_reference_frontend.py can apply the same to the other front ends.
    const double a = R->perturb; const int k = (int)R->snaps.size();
    std::vector<char> used(n, 0), sused(ns, 0);
    for (const Block& B : S.blocks) { used[B.c1] = 1; if (B.c2 >= 0) used[B.c2] = 1; if (B.sw >= 0) sused[B.sw] = 1; }
    for (int i = 0; i < n; ++i) {
      if (!used[i]) continue;
      double* q = R->slam->get_raw_ptr_to_opt_variable_q(i); double* t = R->slam->get_raw_ptr_to_opt_variable_t(i);
      const double e[3] = {a * 0.1 * std::sin(i + k), a * 0.1 * std::cos(2 * i + k), a * 0.1 * std::sin(3 * i + 2 * k)};
      const double nrm = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
      if (nrm > 0) {                                           // ceres::EigenQuaternionParameterization::Plus: [sin|e| e/|e| ; cos|e|] (x) q
        const double sn = std::sin(nrm) / nrm, dx = sn * e[0], dy = sn * e[1], dz = sn * e[2], dw = std::cos(nrm);
        const double x = q[0], y = q[1], z = q[2], w = q[3];
        q[3] = dw * w - dx * x - dy * y - dz * z; q[0] = dw * x + dx * w + dy * z - dz * y; q[1] = dw * y + dy * w + dz * x - dx * z; q[2] = dw * z + dz * w + dx * y - dy * x;
      }
      t[0] += a * std::cos(i + k); t[1] += a * std::sin(2 * i + k); t[2] += a * std::cos(3 * i + k);
    }
    for (int e = 0; e < ns; ++e) if (sused[e]) *R->slam->get_raw_ptr_to_opt_switch(e) = 0.99 - 0.01 * ((7 * e + k) % 50);
  }
}

bool wait_arrival(std::thread& th, long& seen) {   // until that thread is parked in its rate.sleep() again; false after 120 s (a test must fail, not hang)
  ros::Gate& g = ros::gate();
  std::unique_lock<std::mutex> lk(g.m);
  ros::GateState& st = g.per_thread[th.get_id()];
  if (!g.cv.wait_for(lk, std::chrono::seconds(120), [&] { return st.arrivals > seen; })) return false;
  seen = st.arrivals;
  return true;
}
void give_token(std::thread& th) { ros::Gate& g = ros::gate(); std::lock_guard<std::mutex> lk(g.m); ++g.per_thread[th.get_id()].tokens; g.cv.notify_all(); }

}  // namespace

extern "C" {

void* refslam_create() {
  if (g_ref) return nullptr;                      // one instance at a time (the gate is process-wide)
  Ref* R = new Ref();
  R->manager = new NodeDataManager(R->nh);
  R->slam = new PoseGraphSLAM(R->manager);
  ceres::solve_hook() = on_solve;
  R->composer = new Composer(R->manager, R->slam, nullptr, R->nh);   // no VizPoseGraph: only the assembler thread is run
  { ros::Gate& g = ros::gate(); std::lock_guard<std::mutex> lk(g.m); g.per_thread.clear(); g.free_run = false; }
  g_ref = R;
  return R;
}
void refslam_destroy(void* h) {
  Ref* R = (Ref*)h;
  R->slam->reinit_ceres_problem_onnewloopedge_optimize6DOF_disable();
  R->composer->pose_assember_disable();
  { ros::Gate& g = ros::gate(); std::lock_guard<std::mutex> lk(g.m); g.free_run = true; g.cv.notify_all(); }
  if (R->started) R->th.join();
  if (R->composer_started) R->th_composer.join();
  ceres::solve_hook() = nullptr;
  g_ref = nullptr;                                // the reference's destructors free fixed arrays; the objects are leaked on purpose (test process)
  delete R;
}
// NodeDataManager::camera_pose_callback (src/NodeDataManager.cpp:23-103) with a nav_msgs/Odometry built from arrays
void refslam_add_node(void* h, long long stamp_ns, const double* q_xyzw, const double* t) {
  Ref* R = (Ref*)h;
  nav_msgs::Odometry* m = new nav_msgs::Odometry();
  m->header.stamp = ros::Time::fromNSec(stamp_ns);
  m->pose.pose.position.x = t[0]; m->pose.pose.position.y = t[1]; m->pose.pose.position.z = t[2];
  m->pose.pose.orientation.x = q_xyzw[0]; m->pose.pose.orientation.y = q_xyzw[1]; m->pose.pose.orientation.z = q_xyzw[2]; m->pose.pose.orientation.w = q_xyzw[3];
  R->manager->camera_pose_callback(nav_msgs::Odometry::ConstPtr(m));
}
// NodeDataManager::loopclosure_pose_callback (:107-189); returns the number of loop edges the manager holds afterwards
int refslam_add_loop_edge(void* h, long long stamp0_ns, long long stamp1_ns, const double* q_xyzw, const double* t, float weight) {
  Ref* R = (Ref*)h;
  solve_keyframe_pose_graph::LoopEdge* m = new solve_keyframe_pose_graph::LoopEdge();
  m->timestamp0 = ros::Time::fromNSec(stamp0_ns); m->timestamp1 = ros::Time::fromNSec(stamp1_ns);
  m->pose_1T0.position.x = t[0]; m->pose_1T0.position.y = t[1]; m->pose_1T0.position.z = t[2];
  m->pose_1T0.orientation.x = q_xyzw[0]; m->pose_1T0.orientation.y = q_xyzw[1]; m->pose_1T0.orientation.z = q_xyzw[2]; m->pose_1T0.orientation.w = q_xyzw[3];
  m->weight = weight; m->description = "";
  R->manager->loopclosure_pose_callback(solve_keyframe_pose_graph::LoopEdge::ConstPtr(m));
  return R->manager->getEdgeLen();
}
// NodeDataManager::rcvd_kidnap_indicator_callback (:763-792)
void refslam_kidnap(void* h, long long stamp_ns, int kidnapped) {
  Ref* R = (Ref*)h;
  std_msgs::Header* m = new std_msgs::Header();
  m->stamp = ros::Time::fromNSec(stamp_ns); m->frame_id = kidnapped ? "kidnapped" : "unkidnapped";
  R->manager->rcvd_kidnap_indicator_callback(std_msgs::HeaderConstPtr(m));
}
// A solver behind the reference's ceres::Solve call: the callback runs on the reference's solver thread, reads the problem with
// refslam_get_blocks / refslam_get_vars, solves it with whatever it likes and writes the result into the reference's own
// optimisation arrays with refslam_write_vars — which is all ceres::Solve does as far as PoseGraphSLAM.cpp can tell.
void refslam_set_solve_callback(void* h, void (*cb)()) { ((Ref*)h)->solve_callback = cb; }
void refslam_write_vars(void* h, const double* q, const double* t, const double* s) {
  Ref* R = (Ref*)h;
  const int n = R->slam->n_opt_variables(), ns = R->slam->n_opt_switch();
  for (int i = 0; i < n; ++i) { std::memcpy(R->slam->get_raw_ptr_to_opt_variable_q(i), q + 4 * i, 32); std::memcpy(R->slam->get_raw_ptr_to_opt_variable_t(i), t + 3 * i, 24); }
  for (int e = 0; e < ns; ++e) *R->slam->get_raw_ptr_to_opt_switch(e) = s[e];
}
void refslam_set_perturb(void* h, double amplitude) { ((Ref*)h)->perturb = amplitude; }
// One wake-up of PoseGraphSLAM::reinit_ceres_problem_onnewloopedge_optimize6DOF().  Returns 1 if it reached ceres::Solve.
int refslam_wakeup(void* h) {
  Ref* R = (Ref*)h;
  const size_t before = R->snaps.size();
  if (!R->started) {
    R->slam->reinit_ceres_problem_onnewloopedge_optimize6DOF_enable();
    R->th = std::thread(&PoseGraphSLAM::reinit_ceres_problem_onnewloopedge_optimize6DOF, R->slam);
    R->started = true;
  } else give_token(R->th);
  if (!wait_arrival(R->th, R->arrivals_seen)) return -1;
  return R->snaps.size() > before ? 1 : 0;
}
// One pass of Composer::pose_assember_thread (src/Composer.cpp:10-263), the reference's own, on its own thread; then
// global_lmb — the assembled pose of every keyframe — is copied out.  Returns the number of poses.
int refslam_compose_once(void* h, double* T16, int cap) {
  Ref* R = (Ref*)h;
  if (!R->composer_started) {
    R->composer->pose_assember_enable();
    R->th_composer = std::thread(&Composer::pose_assember_thread, R->composer, 30);
    R->composer_started = true;
  } else give_token(R->th_composer);
  if (!wait_arrival(R->th_composer, R->composer_arrivals_seen)) return -1;
  std::lock_guard<std::mutex> lk(R->composer->mx);
  const int n = (int)R->composer->global_lmb.size();
  for (int i = 0; i < n && i < cap; ++i) copy16(R->composer->global_lmb[i], T16 + 16 * i);
  return n;
}
// ---- what the reference built, as of the last ceres::Solve
int refslam_n_blocks(void* h) { Ref* R = (Ref*)h; return R->snaps.empty() ? 0 : (int)R->snaps.back().blocks.size(); }
int refslam_unknown_blocks(void* h) { Ref* R = (Ref*)h; return R->snaps.empty() ? 0 : R->snaps.back().unknown_blocks; }
void refslam_get_blocks(void* h, int* type, int* c1, int* c2, int* sw, double* obs16, double* weight) {
  const Snapshot& S = ((Ref*)h)->snaps.back();
  for (size_t i = 0; i < S.blocks.size(); ++i) { const Block& B = S.blocks[i]; type[i] = B.type; c1[i] = B.c1; c2[i] = B.c2; sw[i] = B.sw; std::memcpy(obs16 + 16 * i, B.obs, 128); weight[i] = B.weight; }
}
int refslam_n_vars(void* h) { Ref* R = (Ref*)h; return R->snaps.empty() ? 0 : (int)R->snaps.back().constant.size(); }
int refslam_n_switches(void* h) { Ref* R = (Ref*)h; return R->snaps.empty() ? 0 : (int)R->snaps.back().s.size(); }
void refslam_get_vars(void* h, double* q, double* t, double* s, int* constant) {
  const Snapshot& S = ((Ref*)h)->snaps.back();
  std::memcpy(q, S.q.data(), 8 * S.q.size()); std::memcpy(t, S.t.data(), 8 * S.t.size());
  if (!S.s.empty()) std::memcpy(s, S.s.data(), 8 * S.s.size());
  std::memcpy(constant, S.constant.data(), 4 * S.constant.size());
}
// ---- live state of the reference objects (call while the solver thread is parked)
int refslam_solved_until(void* h) { return ((Ref*)h)->slam->solvedUntil(); }
int refslam_status(void* h) { return ((Ref*)h)->slam->get_reinit_ceres_problem_onnewloopedge_optimize6DOF_status(); }
int refslam_n_nodes(void* h) { return ((Ref*)h)->manager->getNodeLen(); }
int refslam_n_edges(void* h) { return ((Ref*)h)->manager->getEdgeLen(); }
int refslam_n_worlds(void* h) { return ((Ref*)h)->manager->n_worlds(); }
int refslam_which_world(void* h, long long stamp_ns) { return ((Ref*)h)->manager->which_world_is_this(ros::Time::fromNSec(stamp_ns)); }
int refslam_world_setid(void* h, int w) { return ((Ref*)h)->manager->getWorldsPtr()->find_setID_of_world_i(w); }
int refslam_world_start(void* h, int w) { return ((Ref*)h)->manager->nodeidx_of_world_i_started(w); }
int refslam_world_end(void* h, int w) { return ((Ref*)h)->manager->nodeidx_of_world_i_ended(w); }
int refslam_pose_between_worlds(void* h, int m, int n, double* T16) {      // direct or inverse entries only: the BFS branch of the reference has no return statement
  Worlds* W = ((Ref*)h)->manager->getWorldsPtr();
  if (!W->is_exist(m, n)) return 0;
  copy16(W->getPoseBetweenWorlds(m, n), T16);
  return 1;
}
// ---- the reference's own writers and reader of its state files
// NodeDataManager::saveAsJSON (src/NodeDataManager.cpp:503-628) -> dir/log_posegraph.json, PoseGraphSLAM::saveAsJSON
// (src/PoseGraphSLAM.cpp:1111-1207) -> dir/log_optimized_poses.json, Worlds::saveStateToDisk (src/Worlds.cpp:442-497) -> dir/worlds.json
int refslam_save_json(void* h, const char* dir) {
  Ref* R = (Ref*)h;
  const bool a = R->manager->saveAsJSON(std::string(dir));
  const bool b = R->slam->saveAsJSON(std::string(dir));
  std::ofstream f(std::string(dir) + "/worlds.json");
  f << R->manager->getWorldsPtr()->saveStateToDisk().dump(4);
  return (a ? 1 : 0) | (b ? 2 : 0) | (f.good() ? 4 : 0);
}
// NodeDataManager::loadFromJSON (src/NodeDataManager.cpp:630-754) into this instance's manager (all edges)
int refslam_load_posegraph_json(void* h, const char* dir) { return ((Ref*)h)->manager->loadFromJSON(std::string(dir), std::vector<bool>()) ? 1 : 0; }
// Worlds::loadStateFromDisk (src/Worlds.cpp:499-640) from a file holding the "WorldsData" object
int refslam_load_worlds_json(void* h, const char* file) {
  std::ifstream f(file); if (!f.is_open()) return 0;
  json obj; f >> obj;
  return ((Ref*)h)->manager->getWorldsPtr()->loadStateFromDisk(obj) ? 1 : 0;
}
// The four calls of Composer::loadStateFromDisk (src/Composer.cpp:1137-1167; Composer.cpp itself is ROS visualisation and is not
// compiled here), each of them the reference's own function: 0 on success, else the number of the step that failed.
int refslam_load_state_from_disk(void* h, const char* dir) {
  Ref* R = (Ref*)h;
  std::ifstream f(std::string(dir) + "/solved_posegraph.json"); if (!f.is_open()) return -1;
  json obj; f >> obj;
  if (!R->manager->getWorldsPtr()->loadStateFromDisk(obj["WorldsData"])) return 1;
  if (!R->manager->load_kidnap_data_from_json(obj["KidnapTimestamps"])) return 2;
  if (!R->manager->load_solved_posegraph_data_from_json(obj)) return 3;
  if (!R->slam->load_state()) return 4;
  return 0;
}
// Composer::saveStateToDisk (src/Composer.cpp:990-1105) and Composer::loadStateFromDisk (:1109-1177), the reference's own
int refslam_composer_save(void* h, const char* dir) { return ((Ref*)h)->composer->saveStateToDisk(std::string(dir)) ? 1 : 0; }
int refslam_composer_load(void* h, const char* dir) { return ((Ref*)h)->composer->loadStateFromDisk(std::string(dir)) ? 1 : 0; }
int refslam_slam_n_nodes(void* h) { return ((Ref*)h)->slam->nNodes(); }
int refslam_kidnap_status(void* h) { return ((Ref*)h)->manager->curr_kidnap_status() ? 1 : 0; }
long long refslam_node_stamp(void* h, int i) { return (long long)((Ref*)h)->manager->getNodeTimestamp(i).toNSec(); }
void refslam_manager_node_pose(void* h, int i, double* T16) { copy16(((Ref*)h)->manager->getNodePose(i), T16); }
void refslam_edge(void* h, int e, int* a, int* b, double* T16, double* w) {
  NodeDataManager* m = ((Ref*)h)->manager;
  const std::pair<int, int> p = m->getEdgeIdxInfo(e); *a = p.first; *b = p.second; copy16(m->getEdgePose(e), T16); *w = m->getEdgeWeight(e);
}
int refslam_n_kidnaps(void* h) { return ((Ref*)h)->manager->n_kidnaps(); }
void refslam_get_node_pose(void* h, int i, double* T16) { copy16(((Ref*)h)->slam->getNodePose(i), T16); }

}  // extern "C"

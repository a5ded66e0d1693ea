#include "PoseGraphSLAM.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>
#include <thread>

namespace pgs {

PoseGraphSLAM::PoseGraphSLAM(NodeDataManager* _manager, const PoseGraphSLAMOptions& options) : manager(_manager), opt_(options) {
  solved_until = 0;
  isEnabled = false;
  status = -1;
}

PoseGraphSLAM::~PoseGraphSLAM() { if (handle_) pgs_destroy(handle_); }

// ---------------------------------------------------------------- getters (PoseGraphSLAM.cpp:178-224)
void PoseGraphSLAM::getAllNodePose(std::vector<Matrix4d>& w_T_ci) const {
  w_T_ci.clear();
  const int n = nNodes();
  for (int i = 0; i < n; ++i) w_T_ci.push_back(getNodePose(i));
}
void PoseGraphSLAM::getAllNodeRaw(std::vector<double>& q, std::vector<double>& t) const { std::lock_guard<std::mutex> lk(mutex_opt_vars); q = _opt_quat_; t = _opt_t_; }
int PoseGraphSLAM::nNodes() const { std::lock_guard<std::mutex> lk(mutex_opt_vars); return (int)(_opt_t_.size() / 3); }
const Matrix4d PoseGraphSLAM::getNodePose(int i) const {
  std::lock_guard<std::mutex> lk(mutex_opt_vars);
  if (i < 0 || 3 * (size_t)i >= _opt_t_.size()) return Matrix4d::Identity();   // the reference asserts (compiled out in Release)
  return raw_xyzw_to_mat(&_opt_quat_[4 * (size_t)i], &_opt_t_[3 * (size_t)i]);
}
bool PoseGraphSLAM::nodePoseExists(int i) const { std::lock_guard<std::mutex> lk(mutex_opt_vars); return i >= 0 && 3 * (size_t)i < _opt_t_.size(); }
double PoseGraphSLAM::get_loopedge_switching_variable_val(int i) const {
  std::lock_guard<std::mutex> lk(mutex_opt_vars);
  if (i < 0 || i >= (int)_opt_switch_.size()) return std::numeric_limits<double>::quiet_NaN();   // reference dereferences NULL here
  return _opt_switch_[i];
}
const std::tuple<int, int, float, std::string>& PoseGraphSLAM::get_odomedge_residue_info(int i) const { std::lock_guard<std::mutex> lk(mutex_residue_info); return odometry_edges_terms[i]; }
int PoseGraphSLAM::get_odomedge_residue_info_size() const { std::lock_guard<std::mutex> lk(mutex_residue_info); return (int)odometry_edges_terms.size(); }
const std::tuple<int, int, float, std::string, std::string>& PoseGraphSLAM::get_loopedge_residue_info(int i) const { std::lock_guard<std::mutex> lk(mutex_residue_info); return loop_edges_terms[i]; }
int PoseGraphSLAM::get_loopedge_residue_info_size() const { std::lock_guard<std::mutex> lk(mutex_residue_info); return (int)loop_edges_terms.size(); }

// ---------------------------------------------------------------- variable store (PoseGraphSLAM.cpp:226-361)
int PoseGraphSLAM::n_opt_variables() const { return nNodes(); }
int PoseGraphSLAM::n_opt_switch() const { std::lock_guard<std::mutex> lk(mutex_opt_vars); return (int)_opt_switch_.size(); }
void PoseGraphSLAM::allocate_and_append_new_opt_variable_withpose(const Matrix4d& pose) {
  double q[4], t[3];
  mat_to_raw_xyzw(pose, q, t);
  std::lock_guard<std::mutex> lk(mutex_opt_vars);
  _opt_quat_.insert(_opt_quat_.end(), q, q + 4);
  _opt_t_.insert(_opt_t_.end(), t, t + 3);
}
bool PoseGraphSLAM::update_opt_variable_with(int i, const Matrix4d& pose) {
  double q[4], t[3];
  mat_to_raw_xyzw(pose, q, t);
  std::lock_guard<std::mutex> lk(mutex_opt_vars);
  if (i < 0 || 3 * (size_t)i >= _opt_t_.size()) return false;
  for (int k = 0; k < 4; ++k) _opt_quat_[4 * (size_t)i + k] = q[k];
  for (int k = 0; k < 3; ++k) _opt_t_[3 * (size_t)i + k] = t[k];
  return true;
}
void PoseGraphSLAM::allocate_and_append_new_edge_switch_var() {
  std::lock_guard<std::mutex> lk(mutex_opt_vars);
  _opt_switch_.push_back(opt_.solver.switch_init);   // 0.99, PoseGraphSLAM.cpp:353
}

// ---------------------------------------------------------------- explicit-graph API
bool PoseGraphSLAM::addOdometryEdge(int a, int b, const Matrix4d& a_T_b, double weight) {
  if (a < 0 || b < 0 || a == b || a >= manager->getNodeLen() || b >= manager->getNodeLen()) return false;
  OdomTerm o; o.u = a; o.umf = b; o.weight = weight;
  mat_to_raw_xyzw(a_T_b, o.q, o.t);
  { std::lock_guard<std::mutex> lk(mutex_pending_); pending_explicit_odom_.push_back(o); }
  return true;
}
bool PoseGraphSLAM::addLoopEdge(int a, int b, const Matrix4d& b_T_a, double weight) { return manager->add_loop_edge_by_index(a, b, b_T_a, weight, "addLoopEdge"); }

// ---------------------------------------------------------------- the switched-off builds' blocks
bool PoseGraphSLAM::alternative_terms(int kind, AlternativeTerms& out) const {
  if (kind < 0 || kind > 2) return false;
  out = AlternativeTerms();
  out.kind = kind;
  std::vector<double> q, t, sw;
  { std::lock_guard<std::mutex> lk(mutex_opt_vars); q = _opt_quat_; t = _opt_t_; sw = _opt_switch_; }
  const int n = (int)(t.size() / 3);
  out.n_nodes = n; out.t = t;
  if (kind == 2) {                                       // allocate_and_append_new_opt_variable_withpose under __USE_YPR_REP (:228-247)
    out.rot.resize(3 * (size_t)n);
    for (int i = 0; i < n; ++i) { double tt[3]; mat_to_rawyprt(raw_xyzw_to_mat(&q[4 * (size_t)i], &t[3 * (size_t)i]), &out.rot[3 * (size_t)i], tt); }
  } else out.rot = q;
  auto push_qin = [&](int i, int j, const Matrix4d& i_T_j, int pitch_roll_of) {   // QinFourDOFWeightError::Create(t, yaw(i_T_j), pitch, roll)
    double rel[3], tr[3], own[3], tmp[3];
    mat_to_rawyprt(i_T_j, rel, tr);
    mat_to_rawyprt(manager->getNodePose(pitch_roll_of), own, tmp);
    out.c1.push_back(i); out.c2.push_back(j);
    out.obs_t.insert(out.obs_t.end(), tr, tr + 3);
    out.obs_rot.push_back(rel[0]); out.obs_rot.push_back(own[1]); out.obs_rot.push_back(own[2]);
  };
  if (kind == 0 || kind == 2) {
    for (const OdomTerm& o : odom_terms_) {
      if (kind == 2) { push_qin(o.u, o.umf, raw_xyzw_to_mat(o.q, o.t), o.u); continue; }      // pitch / roll of w_M_u (:1609-1620)
      out.c1.push_back(o.u); out.c2.push_back(o.umf);
      out.obs_rot.insert(out.obs_rot.end(), o.q, o.q + 4); out.obs_t.insert(out.obs_t.end(), o.t, o.t + 3); out.weight.push_back(o.weight);
    }
  }
  if (kind == 1 || kind == 2) {
    for (int e = 0; e < (int)loop_slot_.size(); ++e) {
      if (loop_slot_[e] < 0) continue;                                                            // an endpoint in a dead zone: no block (:1397-1401)
      const std::pair<int, int> paur = manager->getEdgeIdxInfo(e);
      const Matrix4d bTa = manager->getEdgePose(e);
      if (kind == 2) { push_qin(paur.second, paur.first, bTa, paur.first); continue; }          // __w_T_first___ypr[1], [2] (:1389-1392,1541-1543)
      double oq[4], ot[3]; mat_to_raw_xyzw(bTa, oq, ot);
      out.c1.push_back(paur.second); out.c2.push_back(paur.first);
      out.obs_rot.insert(out.obs_rot.end(), oq, oq + 4); out.obs_t.insert(out.obs_t.end(), ot, ot + 3);
      out.weight.push_back(manager->getEdgeWeight(e)); out.sw.push_back(e < (int)sw.size() ? sw[e] : opt_.solver.switch_init);
    }
  }
  return true;
}

// ---------------------------------------------------------------- the solver thread
void PoseGraphSLAM::reinit_ceres_problem_onnewloopedge_optimize6DOF() {
  const auto period = std::chrono::duration<double>(1.0 / std::max(1e-3, opt_.loop_rate_hz));
  while (isEnabled) {
    if (solve_once(false)) ++n_solves_;
    status = 0;
    std::this_thread::sleep_for(period);
  }
}

bool PoseGraphSLAM::load_state() {
  clear_error();
  const int node_len = manager->getNodeLen();
  if (node_len == 0) return fail("load_state: no keyframes in the manager (the reference exits here, PoseGraphSLAM.cpp:54-59)");
  const Worlds* worlds = manager->getWorldsConstPtr();
  for (int yp = n_opt_variables(); yp < node_len; ++yp) {
    const int world = manager->which_world_is_this(manager->getNodeTimestamp(yp));
    const int setid = worlds->find_setID_of_world_i(world);
    Matrix4d ws_T_w = Matrix4d::Identity();
    if (world >= 0 && world != setid) {                                    // :104-116
      bool ok = false;
      if (worlds->is_exist(setid, world)) ws_T_w = worlds->getPoseBetweenWorlds(setid, world, &ok);
      if (!ok) return fail("load_state: no relative pose between world " + std::to_string(world) + " and its set root " + std::to_string(setid));
    }
    allocate_and_append_new_opt_variable_withpose(ws_T_w * manager->getNodePose(yp));   // :118-131
  }
  n_constant_ = node_len;                                                  // SetParameterBlockConstant, :150-151
  { std::lock_guard<std::mutex> lk(mutex_opt_vars); solved_until = node_len - 1; }     // :165
  return true;
}

bool PoseGraphSLAM::solve_once(bool force) {
  clear_error();
  const int node_len = manager->getNodeLen();
  const int loopedge_len = manager->getEdgeLen();
  // trigger only on new loop edges, never while kidnapped (PoseGraphSLAM.cpp:1306-1319)
  bool explicit_pending;
  { std::lock_guard<std::mutex> lk(mutex_pending_); explicit_pending = !pending_explicit_odom_.empty(); }
  if (!force && prev_loopedge_len == loopedge_len && !explicit_pending && !retry_pending_) { status = 0; return false; }
  if (manager->curr_kidnap_status()) { status = 0; return false; }
  if (node_len == 0) { status = 0; return false; }
  status = 1;
  Worlds* worlds = manager->getWorldsPtr();

  // -0- new optimisation variables (identity for now; step 4 writes the guesses)   [:1340-1367]
  for (int yp = n_opt_variables(); yp < node_len; ++yp) allocate_and_append_new_opt_variable_withpose(Matrix4d::Identity());
  for (int yp = n_opt_switch(); yp < loopedge_len; ++yp) allocate_and_append_new_edge_switch_var();

  // -1/2- loop edges, intra and inter world   [:1381-1559]
  loop_slot_.resize(loopedge_len, -1);
  for (int e = loops_taken_until_; e < loopedge_len; ++e) {
    const Matrix4d bTa = manager->getEdgePose(e);
    const double weight = manager->getEdgeWeight(e);
    const std::pair<int, int> paur = manager->getEdgeIdxInfo(e);
    const int _a = paur.first, _b = paur.second;
    if (_a == _b) continue;                     // both stamps resolved to the same keyframe: no constraint, no block
    const int a_world = manager->which_world_is_this(manager->getNodeTimestamp(_a));
    const int b_world = manager->which_world_is_this(manager->getNodeTimestamp(_b));
    if (a_world < 0 || b_world < 0) continue;   // an endpoint lies in a dead zone; its switch slot stays unused
    if (a_world != b_world && !worlds->is_exist(b_world, a_world)) {
      // first edge between two unconnected sets fixes the relative pose of the worlds from ODOMETRY poses [:1459-1464]
      const Matrix4d wa_T_a = manager->getNodePose(_a);
      const Matrix4d wb_T_b = manager->getNodePose(_b);
      const Matrix4d wb_T_wa = (wb_T_b * bTa) * wa_T_a.inverse();
      std::map<int, int> before, after;
      worlds->getWorld2SetIDMap(before);
      worlds->setPoseBetweenWorlds(b_world, a_world, wb_T_wa, "this pose computed from edge " + std::to_string(_a) + " <--> " + std::to_string(_b));
      worlds->getWorld2SetIDMap(after);
      changes_to_setid_on_set_union.clear();   // only the last merge of a trigger survives [:1509]
      for (const auto& kv : before) {
        const int now = after.at(kv.first);
        if (kv.second != now) changes_to_setid_on_set_union[kv.first] = std::make_tuple(kv.second, now);
      }
    }
    double q[4], t[3];
    mat_to_raw_xyzw(bTa, q, t);
    loop_slot_[e] = (int)loop_a_.size();
    loop_a_.push_back(_a); loop_b_.push_back(_b);
    loop_q_.insert(loop_q_.end(), q, q + 4); loop_t_.insert(loop_t_.end(), t, t + 3); loop_w_.push_back(weight);
    { std::lock_guard<std::mutex> lk(mutex_residue_info); loop_edges_terms.push_back(std::make_tuple(_a, _b, (float)weight, std::string(""), std::string(""))); }
  }

  loops_taken_until_ = loopedge_len;

  // -3- odometry edges u <-> u-f for u in (solvedUntil, node_len)   [:1570-1639]
  std::vector<OdomTerm> new_odom;
  if (opt_.derive_odometry) {
    for (int u = std::max(solvedUntil(), odom_added_until_) + 1; u < node_len; ++u) {
      const int world_of_u = manager->which_world_is_this(manager->getNodeTimestamp(u));
      const int set_u = worlds->find_setID_of_world_i(world_of_u);
      for (int f = 1; f <= opt_.odom_fanout; ++f) {
        int world_umf = -1;
        if (u - f >= 0) world_umf = manager->which_world_is_this(manager->getNodeTimestamp(u - f));
        const int set_umf = worlds->find_setID_of_world_i(world_umf);
        if (set_u < 0 || set_umf < 0) continue;   // note: same-world is NOT checked, as in the reference
        if (u - f < 0) continue;
        const Matrix4d w_M_u = manager->getNodePose(u);
        const Matrix4d w_M_umf = manager->getNodePose(u - f);
        const Matrix4d u_M_umf = w_M_u.inverse() * w_M_umf;
        double ypr[3];
        R2ypr(u_M_umf, ypr);   // degrees
        OdomTerm o; o.u = u; o.umf = u - f;
        o.weight = std::pow(opt_.odom_decay, f) * std::exp(-ypr[0] * ypr[0] / opt_.odom_yaw_divisor);
        mat_to_raw_xyzw(u_M_umf, o.q, o.t);
        new_odom.push_back(o);
      }
    }
  }
  { std::lock_guard<std::mutex> lk(mutex_pending_);
    for (const OdomTerm& o : pending_explicit_odom_) new_odom.push_back(o);
    pending_explicit_odom_.clear(); }
  {
    std::lock_guard<std::mutex> lk(mutex_residue_info);
    for (const OdomTerm& o : new_odom) odometry_edges_terms.push_back(std::make_tuple(o.u, o.umf, (float)o.weight, std::string("")));
  }
  odom_terms_.insert(odom_terms_.end(), new_odom.begin(), new_odom.end());
  if (opt_.derive_odometry) odom_added_until_ = std::max(odom_added_until_, node_len - 1);
  retry_pending_ = true;   // cleared when the solve below has gone through

  // -4- initial guesses for every node   [:1649-1793]
  {
    const int su = solvedUntil();
    int su_world = manager->which_world_is_this(manager->getNodeTimestamp(su));
    if (su_world < 0) su_world = -su_world - 1;
    for (int u = 0; u < node_len; ++u) {
      const int world_of_u = manager->which_world_is_this(manager->getNodeTimestamp(u));
      const int set_u = worlds->find_setID_of_world_i(world_of_u);
      if (set_u < 0) continue;   // kidnapped node
      Matrix4d wset_T_w = Matrix4d::Identity();
      if (set_u != world_of_u) {
        bool ok = true;
        if (worlds->is_exist(set_u, world_of_u)) wset_T_w = worlds->getPoseBetweenWorlds(set_u, world_of_u, &ok); else ok = false;
        if (!ok) return fail("initial guess: no pose between set " + std::to_string(set_u) + " and world " + std::to_string(world_of_u) + " (reference exit(3))");
      }
      const bool before = (u <= su);
      const bool in_change = changes_to_setid_on_set_union.count(world_of_u) > 0;
      if (in_change && before) {
        if (set_u == su_world) return fail("initial guess: changed set equals the last solved world (reference exit(8))");
        const int old_setid = std::get<0>(changes_to_setid_on_set_union[world_of_u]);
        const int new_setid = std::get<1>(changes_to_setid_on_set_union[world_of_u]);
        bool ok = true;
        const Matrix4d wsetnew_T_wsetold = worlds->getPoseBetweenWorlds(new_setid, old_setid, &ok);
        if (!ok) return fail("initial guess: no pose between new and old set");
        update_opt_variable_with(u, wsetnew_T_wsetold * this->getNodePose(u));
      } else if (!before) {
        // both !before branches of the reference are identical [:1727-1751, :1766-1786]
        if (su_world == world_of_u) {
          const Matrix4d last_M_u = manager->getNodePose(su).inverse() * manager->getNodePose(u);
          update_opt_variable_with(u, this->getNodePose(su) * last_M_u);     // dead-reckon from the last solved node
        } else {
          update_opt_variable_with(u, wset_T_w * manager->getNodePose(u));   // odometry pose mapped into the set root's frame
        }
      } else if (su == 0) {
        update_opt_variable_with(u, manager->getNodePose(u));                // very first trigger [:1756-1760]
      }                                                                      // else: keep the previous solution
    }
  }

  // -5- gauge by regularisation: one soft anchor on the first node of every set-root world   [:1801-1879]
  reg_terms_.clear();
  for (int ww = 0; ww < manager->n_worlds(); ++ww) {
    const int ww_setid = worlds->find_setID_of_world_i(ww);
    const int ww_start = manager->nodeidx_of_world_i_started(ww);
    const int ww_end = manager->nodeidx_of_world_i_ended(ww);
    if (ww_start < 0) continue;
    if (ww_setid >= 0 && ww_setid == ww) {
      // std::max(1.1, x) keeps 1.1 when x is NaN (ww_end = -1 makes the log argument non-positive)
      const double x = std::log((double)(1 + ww_end - ww_start)) / 2.0;
      RegTerm r; r.node = ww_start; r.anchor = this->getNodePose(ww_start);   // anchored at the CURRENT estimate [:1844]
      r.weight = (opt_.min_reg_weight < x) ? x : opt_.min_reg_weight;
      reg_terms_.push_back(r);
    }
  }
  changes_to_setid_on_set_union.clear();

  // -6- hand the problem to the device and solve   [:1887-1924]
  if (!opt_.dry_run) {
    if (!handle_) {
      if (pgs_create(&opt_.solver, &handle_) != PGS_OK) return fail(std::string("pgs_create: ") + pgs_last_error(nullptr));
    }
    std::vector<double> q, t;
    { std::lock_guard<std::mutex> lk(mutex_opt_vars); q = _opt_quat_; t = _opt_t_; }
    int rc = PGS_OK;
    // every counter moves as soon as its call has succeeded: what a failed trigger already put on the device is not sent again
    if (node_len > n_device_nodes_) {
      rc = pgs_append_nodes(handle_, node_len - n_device_nodes_, &q[4 * (size_t)n_device_nodes_], &t[3 * (size_t)n_device_nodes_]);
      if (rc == PGS_OK) n_device_nodes_ = node_len;
    }
    if (rc == PGS_OK) rc = pgs_update_nodes(handle_, 0, node_len, q.data(), t.data());
    if (rc == PGS_OK && n_constant_ > n_constant_on_device_) { rc = pgs_set_constant_nodes(handle_, 0, n_constant_, 1); if (rc == PGS_OK) n_constant_on_device_ = n_constant_; }
    const int n_loops = (int)loop_a_.size();
    if (rc == PGS_OK && n_loops > n_device_loops_) {
      const size_t f = (size_t)n_device_loops_;
      rc = pgs_add_loop_edges(handle_, n_loops - n_device_loops_, &loop_a_[f], &loop_b_[f], &loop_q_[4 * f], &loop_t_[3 * f], &loop_w_[f]);
      if (rc == PGS_OK) n_device_loops_ = n_loops;
    }
    if (rc == PGS_OK && (int)odom_terms_.size() > n_device_odom_) {
      std::vector<int> c1, c2; std::vector<double> oq, ot, ow;
      for (size_t k = (size_t)n_device_odom_; k < odom_terms_.size(); ++k) {
        const OdomTerm& o = odom_terms_[k];
        c1.push_back(o.u); c2.push_back(o.umf); oq.insert(oq.end(), o.q, o.q + 4); ot.insert(ot.end(), o.t, o.t + 3); ow.push_back(o.weight);
      }
      rc = pgs_add_odom_edges(handle_, (int)c1.size(), c1.data(), c2.data(), oq.data(), ot.data(), ow.data());
      if (rc == PGS_OK) n_device_odom_ = (int)odom_terms_.size();
    }
    if (rc == PGS_OK) {
      std::vector<int> rn; std::vector<double> rq, rt, rw;
      for (const RegTerm& r : reg_terms_) { double qq[4], tt[3]; mat_to_raw_xyzw(r.anchor, qq, tt); rn.push_back(r.node); rq.insert(rq.end(), qq, qq + 4); rt.insert(rt.end(), tt, tt + 3); rw.push_back(r.weight); }
      rc = pgs_set_regularizers(handle_, (int)rn.size(), rn.data(), rq.data(), rt.data(), rw.data());
    }
    if (rc != PGS_OK) return fail(std::string("device problem update: ") + pgs_last_error(handle_));

    status = 2;
    pgs_summary sum{};
    std::vector<pgs_iteration> its(opt_.solver.max_num_iterations + 8, pgs_iteration{});
    rc = pgs_solve(handle_, &sum, its.data(), (int)its.size());
    if (rc != PGS_OK) return fail(std::string("pgs_solve: ") + pgs_last_error(handle_));
    its.resize(std::min<size_t>(its.size(), (size_t)std::max(sum.num_iterations, 0)));
    { std::lock_guard<std::mutex> lk(mutex_summary_); summary_ = sum; iterations_.swap(its); }   // readers on other threads get copies
    // read the solution back and publish it under the mutex in one go: like Ceres with
    // update_state_every_iteration=false, readers only ever see the state of a finished solve
    std::vector<double> sw(n_device_loops_);
    rc = pgs_get_poses(handle_, 0, node_len, q.data(), t.data());
    if (rc == PGS_OK && n_device_loops_) rc = pgs_get_switches(handle_, 0, n_device_loops_, sw.data());
    if (rc != PGS_OK) return fail(std::string("read-back: ") + pgs_last_error(handle_));
    {
      std::lock_guard<std::mutex> lk(mutex_opt_vars);
      _opt_quat_ = q; _opt_t_ = t;
      for (int e = 0; e < loopedge_len; ++e) if (loop_slot_[e] >= 0) _opt_switch_[e] = sw[loop_slot_[e]];
    }
  } else {
    n_device_loops_ = (int)loop_a_.size();
  }
  retry_pending_ = false;
  { std::lock_guard<std::mutex> lk(mutex_opt_vars); solved_until = node_len - 1; }   // unconditionally [:1908]
  status = 3;
  prev_loopedge_len = loopedge_len;
  prev_node_len = node_len;
  return true;
}

}  // namespace pgs

#include "PoseGraphSLAM.h"
#include "host_lap.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>
#include <cstdlib>
#include <thread>

namespace pgs {

namespace {
// fn(chunk, lo, hi) over [begin, end) in contiguous chunks, one host thread per chunk; the chunks are in index order, so
// results gathered chunk by chunk come out exactly as a sequential loop would produce them
template <class F>
void for_chunks(int begin, int end, int min_per_chunk, std::vector<int>* bounds, F&& fn) {
  const int n = std::max(0, end - begin);
  int T = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (const char* e = getenv("PGS_HOST_THREADS")) T = std::max(1, std::min(16, atoi(e)));
  T = std::max(1, std::min(T, n / std::max(1, min_per_chunk)));
  bounds->resize(T + 1);
  for (int k = 0; k <= T; ++k) (*bounds)[k] = begin + (int)((long long)n * k / T);
  if (T == 1) { fn(0, begin, end); return; }
  std::vector<std::thread> th;
  for (int k = 1; k < T; ++k) th.emplace_back([&, k] { fn(k, (*bounds)[k], (*bounds)[k + 1]); });
  fn(0, (*bounds)[0], (*bounds)[1]);
  for (std::thread& t : th) t.join();
}
}  // namespace

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
  int loopedge_len = manager->getEdgeLen();
  // threaded ingest: a loop edge that arrived after node_len was read may name a keyframe this trigger does not know yet;
  // it and the edges behind it wait for the next trigger
  for (int e = loops_taken_until_; e < loopedge_len; ++e) {
    const std::pair<int, int> ab = manager->getEdgeIdxInfo(e);
    if (ab.first >= node_len || ab.second >= node_len) { loopedge_len = e; break; }
  }
  // trigger only on new loop edges, never while kidnapped (PoseGraphSLAM.cpp:1306-1319)
  bool explicit_pending;
  { std::lock_guard<std::mutex> lk(mutex_pending_); explicit_pending = !pending_explicit_odom_.empty(); }
  if (!force && prev_loopedge_len == loopedge_len && !explicit_pending && !retry_pending_) { status = 0; return false; }
  if (manager->curr_kidnap_status()) { status = 0; return false; }
  if (node_len == 0) { status = 0; return false; }
  status = 1;
  Worlds* worlds = manager->getWorldsPtr();
  HostLap lap;

  // -0- new optimisation variables (identity for now; step 4 writes the guesses)   [:1340-1367]
  {
    double qi[4], ti[3];
    mat_to_raw_xyzw(Matrix4d::Identity(), qi, ti);
    std::lock_guard<std::mutex> lk(mutex_opt_vars);
    for (int yp = (int)(_opt_t_.size() / 3); yp < node_len; ++yp) { _opt_quat_.insert(_opt_quat_.end(), qi, qi + 4); _opt_t_.insert(_opt_t_.end(), ti, ti + 3); }
    if ((int)_opt_switch_.size() < loopedge_len) _opt_switch_.resize(loopedge_len, opt_.solver.switch_init);   // 0.99, PoseGraphSLAM.cpp:353
  }

  // odometry pose and world of every keyframe: one locked snapshot instead of a lock per call
  std::vector<Matrix4d> m_pose; std::vector<int> world_of;
  manager->snapshot_nodes(0, node_len, m_pose, world_of);

  // -1/2- loop edges, intra and inter world   [:1381-1559]
  loop_slot_.resize(loopedge_len, -1);
  { std::lock_guard<std::mutex> lk(mutex_residue_info); loop_edges_terms.reserve(loop_edges_terms.size() + (size_t)std::max(0, loopedge_len - loops_taken_until_)); }
  for (int e = loops_taken_until_; e < loopedge_len; ++e) {
    const Matrix4d bTa = manager->getEdgePose(e);
    const double weight = manager->getEdgeWeight(e);
    const std::pair<int, int> paur = manager->getEdgeIdxInfo(e);
    const int _a = paur.first, _b = paur.second;
    if (_a == _b) continue;                     // both stamps resolved to the same keyframe: no constraint, no block
    const int a_world = world_of[_a], b_world = world_of[_b];              // which_world_is_this(getNodeTimestamp(.))
    if (a_world < 0 || b_world < 0) continue;   // an endpoint lies in a dead zone; its switch slot stays unused
    if (a_world != b_world && !worlds->is_exist(b_world, a_world)) {
      // first edge between two unconnected sets fixes the relative pose of the worlds from ODOMETRY poses [:1459-1464]
      const Matrix4d& wa_T_a = m_pose[_a];
      const Matrix4d& wb_T_b = m_pose[_b];
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

  // The per-keyframe rules below read nothing but the snapshot and the set of a world, so the loops run in contiguous chunks on
  // the host's cores (every keyframe's terms depend on that keyframe alone; the chunks are concatenated in order, so the lists
  // are those of the sequential loop).  PGS_HOST_THREADS caps the thread count (default: the host's cores, at most 16).
  std::vector<int> set_of_world;
  { std::map<int, int> w2s; worlds->getWorld2SetIDMap(w2s); for (const auto& kv : w2s) { if (kv.first >= (int)set_of_world.size()) set_of_world.resize(kv.first + 1, -1); set_of_world[kv.first] = kv.second; } }
  auto set_of = [&](int world) { return (world >= 0 && world < (int)set_of_world.size()) ? set_of_world[world] : -1; };   // find_setID_of_world_i

  // -3- odometry edges u <-> u-f for u in (solvedUntil, node_len)   [:1570-1639]
  std::vector<OdomTerm> new_odom;
  if (opt_.derive_odometry) {
    const int u0 = std::max(solvedUntil(), odom_added_until_) + 1;
    std::vector<int> bounds; std::vector<std::vector<OdomTerm>> part(16);
    for_chunks(u0, node_len, 2048, &bounds, [&](int k, int lo, int hi) {
      std::vector<OdomTerm>& out = part[k];
      out.reserve((size_t)std::max(0, hi - lo) * (size_t)std::max(1, opt_.odom_fanout));
      for (int u = lo; u < hi; ++u) {
        const int set_u = set_of(world_of[u]);
        if (set_u < 0) continue;
        const Matrix4d u_M_w = m_pose[u].inverse();
        for (int f = 1; f <= opt_.odom_fanout; ++f) {
          if (u - f < 0) continue;
          if (set_of(world_of[u - f]) < 0) continue;   // note: same-world is NOT checked, as in the reference
          const Matrix4d u_M_umf = u_M_w * m_pose[u - f];
          double ypr[3];
          R2ypr(u_M_umf, ypr);   // degrees
          OdomTerm o; o.u = u; o.umf = u - f;
          o.weight = std::pow(opt_.odom_decay, f) * std::exp(-ypr[0] * ypr[0] / opt_.odom_yaw_divisor);
          mat_to_raw_xyzw(u_M_umf, o.q, o.t);
          out.push_back(o);
        }
      }
    });
    for (size_t k = 0; k + 1 < bounds.size(); ++k) new_odom.insert(new_odom.end(), part[k].begin(), part[k].end());
  }
  { std::lock_guard<std::mutex> lk(mutex_pending_);
    for (const OdomTerm& o : pending_explicit_odom_) new_odom.push_back(o);
    pending_explicit_odom_.clear(); }
  {
    std::lock_guard<std::mutex> lk(mutex_residue_info);
    odometry_edges_terms.reserve(odometry_edges_terms.size() + new_odom.size());
    for (const OdomTerm& o : new_odom) odometry_edges_terms.emplace_back(o.u, o.umf, (float)o.weight, std::string());
  }
  odom_terms_.insert(odom_terms_.end(), new_odom.begin(), new_odom.end());
  if (opt_.derive_odometry) odom_added_until_ = std::max(odom_added_until_, node_len - 1);
  retry_pending_ = true;   // cleared when the solve below has gone through

  // -4- initial guesses for every node   [:1649-1793]
  {
    const int su = solvedUntil();
    int su_world = (su >= 0 && su < node_len) ? world_of[su] : manager->which_world_is_this(manager->getNodeTimestamp(su));
    if (su_world < 0) su_world = -su_world - 1;
    // relative poses between a world and its set root, and between the sets a merge of this trigger changed: looked up once
    std::vector<Matrix4d> wset_T_w_of(set_of_world.size(), Matrix4d::Identity());
    std::vector<char> wset_ok(set_of_world.size(), 1);
    for (int w = 0; w < (int)set_of_world.size(); ++w) {
      const int set_w = set_of_world[w];
      if (set_w < 0 || set_w == w) continue;
      bool ok = true;
      if (worlds->is_exist(set_w, w)) wset_T_w_of[w] = worlds->getPoseBetweenWorlds(set_w, w, &ok); else ok = false;
      wset_ok[w] = ok ? 1 : 0;
    }
    std::map<int, Matrix4d> wsetnew_T_wsetold_of; std::map<int, char> change_ok;
    for (const auto& kv : changes_to_setid_on_set_union) {
      bool ok = true;
      wsetnew_T_wsetold_of[kv.first] = worlds->getPoseBetweenWorlds(std::get<1>(kv.second), std::get<0>(kv.second), &ok);
      change_ok[kv.first] = ok ? 1 : 0;
    }
    std::vector<double> q, t;
    { std::lock_guard<std::mutex> lk(mutex_opt_vars); q = _opt_quat_; t = _opt_t_; }
    const Matrix4d su_M_w = (su >= 0 && su < node_len) ? m_pose[su].inverse() : Matrix4d::Identity();
    Matrix4d opt_su = Matrix4d::Identity();                                         // this->getNodePose(su) as the keyframes after su see it
    auto put = [&](int u, const Matrix4d& pose) { mat_to_raw_xyzw(pose, &q[4 * (size_t)u], &t[3 * (size_t)u]); };   // update_opt_variable_with
    auto guess = [&](int u) -> int {                                                // 0, or why the reference would have exited
      const int world_of_u = world_of[u];
      const int set_u = set_of(world_of_u);
      if (set_u < 0) return 0;   // kidnapped node
      if (set_u != world_of_u && !wset_ok[world_of_u]) return 1;
      const bool before = (u <= su);
      const bool in_change = changes_to_setid_on_set_union.count(world_of_u) > 0;
      if (in_change && before) {
        if (set_u == su_world) return 2;
        if (!change_ok.at(world_of_u)) return 3;
        put(u, wsetnew_T_wsetold_of.at(world_of_u) * raw_xyzw_to_mat(&q[4 * (size_t)u], &t[3 * (size_t)u]));
      } else if (!before) {
        // both !before branches of the reference are identical [:1727-1751, :1766-1786]
        if (su_world == world_of_u) put(u, opt_su * (su_M_w * m_pose[u]));         // dead-reckon from the last solved node
        else put(u, wset_T_w_of[world_of_u] * m_pose[u]);                           // odometry pose mapped into the set root's frame
      } else if (su == 0) {
        put(u, m_pose[u]);                                                          // very first trigger [:1756-1760]
      }                                                                             // else: keep the previous solution
      return 0;
    };
    // Keyframe su first: the keyframes after it dead-reckon from its pose as this very loop leaves it (re-based by a set
    // merge, or the odometry pose on the very first trigger); every other keyframe reads only its own entry.
    int su_why = 0;
    if (su >= 0 && su < node_len) {
      su_why = guess(su);
      opt_su = raw_xyzw_to_mat(&q[4 * (size_t)su], &t[3 * (size_t)su]);
    }
    // first failing keyframe of every chunk (the sequential loop stops at the first one overall)
    std::vector<int> bounds; std::vector<int> bad_u(16, -1), bad_why(16, 0);
    for_chunks(0, node_len, 4096, &bounds, [&](int k, int lo, int hi) {
      for (int u = lo; u < hi; ++u) {
        const int why_u = (u == su) ? su_why : guess(u);
        if (why_u) { bad_u[k] = u; bad_why[k] = why_u; return; }
      }
    });
    int first_bad = -1, why = 0;
    for (size_t k = 0; k + 1 < bounds.size(); ++k) if (bad_u[k] >= 0) { first_bad = bad_u[k]; why = bad_why[k]; break; }
    // keyframes before the first failing one have their guesses, as after the sequential loop's early return
    {
      std::lock_guard<std::mutex> lk(mutex_opt_vars);
      const size_t upto = first_bad >= 0 ? (size_t)first_bad : (size_t)node_len;
      std::copy(q.begin(), q.begin() + 4 * upto, _opt_quat_.begin());
      std::copy(t.begin(), t.begin() + 3 * upto, _opt_t_.begin());
    }
    if (first_bad >= 0) {
      const int world_of_u = world_of[first_bad];
      if (why == 1) return fail("initial guess: no pose between set " + std::to_string(set_of(world_of_u)) + " and world " + std::to_string(world_of_u) + " (reference exit(3))");
      if (why == 2) return fail("initial guess: changed set equals the last solved world (reference exit(8))");
      return fail("initial guess: no pose between new and old set");
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

  lap.lap("trigger rules (host)");
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

    lap.lap("problem -> C-ABI");
    status = 2;
    pgs_summary sum{};
    std::vector<pgs_iteration> its(opt_.solver.max_num_iterations + 8, pgs_iteration{});
    rc = pgs_solve(handle_, &sum, its.data(), (int)its.size());
    if (rc != PGS_OK) return fail(std::string("pgs_solve: ") + pgs_last_error(handle_));
    lap.lap("pgs_solve");
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
    lap.lap("poses + switches back");
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

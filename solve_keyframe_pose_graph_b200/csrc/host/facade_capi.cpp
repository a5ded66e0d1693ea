// C view of the facade (include/pgs_facade.h).
#include "../../../include/pgs_facade.h"
#include "../../../include/pgs_fourdof.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "Composer.h"
#include "GraphIO.h"
#include "PoseGraphSLAM.h"
#include "RosShim.h"

struct pgs_facade_s {
  pgs::NodeDataManager manager;
  pgs::PoseGraphSLAM* slam = nullptr;
  pgs::Composer* composer = nullptr;
  std::thread solver_thread;
  int device = 0;
  std::string err;
  void stop_thread() { if (solver_thread.joinable()) { slam->reinit_ceres_problem_onnewloopedge_optimize6DOF_disable(); solver_thread.join(); } }
  ~pgs_facade_s() { stop_thread(); delete composer; delete slam; }
};

// No C++ exception may cross the C boundary: every multi-statement entry point is a function-try-block.
#define CATCH_FACADE(h)                                                                                                      \
  catch (const std::bad_alloc&) { if (h) (h)->err = "out of host memory"; return PGS_ERR_OUT_OF_MEMORY; }                    \
  catch (const std::exception& e) { if (h) (h)->err = std::string("unexpected C++ exception: ") + e.what(); return PGS_ERR_STATE; } \
  catch (...) { if (h) (h)->err = "unexpected C++ exception"; return PGS_ERR_STATE; }
#define CATCH_IO                                                                                                             \
  catch (const std::bad_alloc&) { return PGS_ERR_OUT_OF_MEMORY; }                                                            \
  catch (...) { return PGS_ERR_STATE; }

extern "C" {

int pgs_facade_default_options(pgs_facade_options* o) try {
  if (!o) return PGS_ERR_INVALID_ARGUMENT;
  o->odom_fanout = 5; o->derive_odometry = 1; o->dry_run = 0;
  return pgs_default_options(&o->solver);
} CATCH_IO
int pgs_facade_create(const pgs_facade_options* o, pgs_facade_handle* out) try {
  if (!out) return PGS_ERR_INVALID_ARGUMENT;
  pgs_facade_options d;
  if (o) d = *o; else pgs_facade_default_options(&d);
  pgs_facade_s* h = new pgs_facade_s();
  pgs::PoseGraphSLAMOptions po;
  po.odom_fanout = d.odom_fanout; po.derive_odometry = d.derive_odometry != 0; po.dry_run = d.dry_run != 0; po.solver = d.solver;
  h->slam = new pgs::PoseGraphSLAM(&h->manager, po);
  h->device = d.solver.device;
  *out = h;
  return PGS_OK;
} CATCH_IO
void pgs_facade_destroy(pgs_facade_handle h) { delete h; }
const char* pgs_facade_last_error(pgs_facade_handle h) {
  static thread_local std::string snapshot;     // valid until the calling thread asks again
  try { snapshot = h ? (h->err.empty() ? h->slam->last_error() : h->err) : std::string(); } catch (...) { return "out of host memory"; }
  return snapshot.c_str();
}

int pgs_facade_add_nodes(pgs_facade_handle h, int32_t n, const int64_t* stamps, const double* q, const double* t) try {
  if (!h || n < 0 || (n && (!stamps || !q || !t))) return PGS_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < n; ++i) h->manager.add_node(stamps[i], pgs::raw_xyzw_to_mat(q + 4 * (size_t)i, t + 3 * (size_t)i));
  return PGS_OK;
} CATCH_FACADE(h)
int pgs_facade_add_loop_edges(pgs_facade_handle h, int32_t m, const int32_t* a, const int32_t* b, const double* q, const double* t, const double* w) try {
  if (!h || m < 0 || (m && (!a || !b || !q || !t))) return PGS_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < m; ++i)
    if (!h->manager.add_loop_edge_by_index(a[i], b[i], pgs::raw_xyzw_to_mat(q + 4 * (size_t)i, t + 3 * (size_t)i), w ? w[i] : 1.0)) {
      h->err = "loop edge endpoint out of range"; return PGS_ERR_INVALID_ARGUMENT; }
  return PGS_OK;
} CATCH_FACADE(h)
int pgs_facade_add_loop_edge_stamped(pgs_facade_handle h, int64_t sa, int64_t sb, const double* q, const double* t, double w) try {
  if (!h || !q || !t) return PGS_ERR_INVALID_ARGUMENT;
  return h->manager.add_loop_edge(sa, sb, pgs::raw_xyzw_to_mat(q, t), w) ? 1 : 0;
} CATCH_FACADE(h)
int pgs_facade_kidnap_indicator(pgs_facade_handle h, int64_t stamp, int32_t kidnapped) try {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  return h->manager.rcvd_kidnap_indicator(stamp, kidnapped != 0) ? PGS_OK : PGS_ERR_STATE;
} CATCH_FACADE(h)
int pgs_facade_add_odometry_edge(pgs_facade_handle h, int32_t a, int32_t b, const double* q, const double* t, double w) try {
  if (!h || !q || !t) return PGS_ERR_INVALID_ARGUMENT;
  return h->slam->addOdometryEdge(a, b, pgs::raw_xyzw_to_mat(q, t), w) ? PGS_OK : PGS_ERR_INVALID_ARGUMENT;
} CATCH_FACADE(h)
static pgs::ros_shim::Pose make_pose(const double* p, const double* q) {
  pgs::ros_shim::Pose P; P.position.x = p[0]; P.position.y = p[1]; P.position.z = p[2];
  P.orientation.x = q[0]; P.orientation.y = q[1]; P.orientation.z = q[2]; P.orientation.w = q[3];
  return P;
}
int pgs_facade_camera_pose_callback(pgs_facade_handle h, uint32_t sec, uint32_t nsec, const double* p, const double* q, const double* cov36) try {
  if (!h || !p || !q) return PGS_ERR_INVALID_ARGUMENT;
  pgs::ros_shim::Odometry msg; msg.header.stamp.sec = sec; msg.header.stamp.nsec = nsec; msg.pose.pose = make_pose(p, q);
  if (cov36) std::memcpy(msg.pose.covariance, cov36, sizeof(double) * 36);
  pgs::ros_shim::camera_pose_callback(h->manager, msg);
  return PGS_OK;
} CATCH_FACADE(h)
int pgs_facade_loopclosure_pose_callback(pgs_facade_handle h, uint32_t sec0, uint32_t nsec0, uint32_t sec1, uint32_t nsec1, const double* p, const double* q, float weight,
                                         const char* description) try {
  if (!h || !p || !q) return PGS_ERR_INVALID_ARGUMENT;
  pgs::ros_shim::LoopEdge msg; msg.timestamp0.sec = sec0; msg.timestamp0.nsec = nsec0; msg.timestamp1.sec = sec1; msg.timestamp1.nsec = nsec1;
  msg.pose_1T0 = make_pose(p, q); msg.weight = weight; msg.description = description ? description : "";
  return pgs::ros_shim::loopclosure_pose_callback(h->manager, msg) ? 1 : 0;
} CATCH_FACADE(h)
int pgs_facade_rcvd_kidnap_indicator_callback(pgs_facade_handle h, uint32_t sec, uint32_t nsec, const char* frame_id) try {
  if (!h || !frame_id) return PGS_ERR_INVALID_ARGUMENT;
  pgs::ros_shim::Header hd; hd.stamp.sec = sec; hd.stamp.nsec = nsec; hd.frame_id = frame_id;
  return pgs::ros_shim::rcvd_kidnap_indicator_callback(h->manager, hd) ? PGS_OK : PGS_ERR_STATE;
} CATCH_FACADE(h)
int pgs_facade_load_state(pgs_facade_handle h) try {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  h->err.clear();
  return h->slam->load_state() ? PGS_OK : PGS_ERR_STATE;
} CATCH_FACADE(h)
int pgs_facade_solve_once(pgs_facade_handle h, int32_t force) try {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  h->err.clear();
  const bool ok = h->slam->solve_once(force != 0);
  if (!ok && !h->slam->last_error().empty()) return PGS_ERR_STATE;
  return ok ? 1 : 0;
} CATCH_FACADE(h)
int pgs_facade_thread_start(pgs_facade_handle h, double rate_hz) try {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  if (h->solver_thread.joinable()) return PGS_ERR_STATE;
  h->slam->set_loop_rate_hz(rate_hz);
  h->slam->reinit_ceres_problem_onnewloopedge_optimize6DOF_enable();
  h->solver_thread = std::thread(&pgs::PoseGraphSLAM::reinit_ceres_problem_onnewloopedge_optimize6DOF, h->slam);
  return PGS_OK;
} CATCH_FACADE(h)
int pgs_facade_thread_stop(pgs_facade_handle h) try {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  h->stop_thread();
  return h->slam->n_solves();
} CATCH_FACADE(h)
int pgs_facade_status(pgs_facade_handle h) { return h ? h->slam->get_reinit_ceres_problem_onnewloopedge_optimize6DOF_status() : -1; }
int32_t pgs_facade_n_nodes(pgs_facade_handle h) { return h ? h->slam->nNodes() : 0; }
int32_t pgs_facade_solved_until(pgs_facade_handle h) { return h ? h->slam->solvedUntil() : 0; }
int pgs_facade_get_poses(pgs_facade_handle h, int32_t cap, double* q, double* t) try {
  if (!h || cap < 0) return PGS_ERR_INVALID_ARGUMENT;
  std::vector<double> qq, tt;
  h->slam->getAllNodeRaw(qq, tt);                       // one consistent snapshot under the variables' mutex
  const int n = std::min<int>((int)(tt.size() / 3), cap);
  if (q && n) std::memcpy(q, qq.data(), sizeof(double) * 4 * (size_t)n);
  if (t && n) std::memcpy(t, tt.data(), sizeof(double) * 3 * (size_t)n);
  return n;
} CATCH_FACADE(h)
int pgs_facade_get_switches(pgs_facade_handle h, int32_t n, double* s) try {
  if (!h || !s) return PGS_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < n; ++i) s[i] = h->slam->get_loopedge_switching_variable_val(i);
  return PGS_OK;
} CATCH_FACADE(h)
int pgs_facade_get_summary(pgs_facade_handle h, pgs_summary* s, pgs_iteration* iters, int32_t cap) try {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  if (s) *s = h->slam->last_summary();
  const std::vector<pgs_iteration> it = h->slam->last_iterations();
  for (int i = 0; iters && i < cap && i < (int)it.size(); ++i) iters[i] = it[i];
  return PGS_OK;
} CATCH_FACADE(h)
int pgs_facade_compose(pgs_facade_handle h, int32_t cap, double* out_T, int32_t* out_world) try {
  if (!h || cap < 0) return PGS_ERR_INVALID_ARGUMENT;
  h->err.clear();
  if (!h->composer) h->composer = new pgs::Composer(&h->manager, h->slam, h->device);
  if (!h->composer->pose_assember_once()) { h->err = h->composer->last_error(); return PGS_ERR_CUDA; }
  const std::vector<pgs::Matrix4d> lmb = h->composer->get_global_lmb();
  const int n = std::min<int>((int)lmb.size(), cap);     // keyframes may have arrived since the caller sized its buffers
  for (int i = 0; i < n; ++i) {
    if (out_T) std::memcpy(out_T + 16 * (size_t)i, lmb[i].m, 128);
    if (out_world) out_world[i] = h->manager.which_world_is_this(h->manager.getNodeTimestamp(i));
  }
  return n;
} CATCH_FACADE(h)
int32_t pgs_facade_n_keyframes(pgs_facade_handle h) { return h ? h->manager.getNodeLen() : 0; }
int pgs_facade_last_known_camerapose(pgs_facade_handle h, double* T16, int64_t* stamp_ns) try {
  if (!h || !h->composer) return -1;
  pgs::Matrix4d T; int64_t st = 0;
  const int r = h->composer->get_last_known_camerapose(T, st);
  if (r >= 0) { if (T16) std::memcpy(T16, T.m, 128); if (stamp_ns) *stamp_ns = st; }
  return r;
} CATCH_FACADE(h)
int pgs_facade_compose_timing(pgs_facade_handle h, double* ms_kernel, double* ms_total) try {
  if (!h || !h->composer) return PGS_ERR_STATE;
  if (ms_kernel) *ms_kernel = h->composer->last_kernel_ms();
  if (ms_total) *ms_total = h->composer->last_total_ms();
  return PGS_OK;
} CATCH_FACADE(h)
int pgs_facade_save_json(pgs_facade_handle h, const char* dir) try {
  if (!h || !dir) return PGS_ERR_INVALID_ARGUMENT;
  h->err.clear();
  int mask = 0;
  if (!pgs::saveAsJSON(h->manager, dir, &h->err)) return PGS_ERR_STATE;
  mask |= 1;
  if (!pgs::saveAsJSON(*h->slam, h->manager, dir, &h->err)) return PGS_ERR_STATE;
  mask |= 2;
  if (!pgs::saveSolvedPoseGraph(h->composer, h->manager, dir, &h->err)) return PGS_ERR_STATE;   // always written: KidnapTimestamps + WorldsData
  if (h->composer && !h->composer->get_global_lmb().empty()) mask |= 4;                          // ... with the assembled poses when a pass has run
  return mask;
} CATCH_FACADE(h)
int pgs_facade_save_state_to_disk(pgs_facade_handle h, const char* dir) try {
  if (!h || !dir) return PGS_ERR_INVALID_ARGUMENT;
  h->err.clear();
  if (!h->manager.curr_kidnap_status()) h->manager.mark_as_kidnapped_and_signal_end_of_world();          // Composer.cpp:969-971
  return pgs::saveSolvedPoseGraph(h->composer, h->manager, dir, &h->err) ? PGS_OK : PGS_ERR_STATE;
} CATCH_FACADE(h)
int pgs_facade_load_posegraph_json(pgs_facade_handle h, const char* dir) try {
  if (!h || !dir) return PGS_ERR_INVALID_ARGUMENT;
  h->err.clear();
  return pgs::loadFromJSON(h->manager, dir, {}, true, &h->err) ? PGS_OK : PGS_ERR_STATE;
} CATCH_FACADE(h)
int pgs_facade_load_worlds_state(pgs_facade_handle h, const char* file) try {
  if (!h || !file) return PGS_ERR_INVALID_ARGUMENT;
  h->err.clear();
  FILE* f = fopen(file, "rb");
  if (!f) { h->err = std::string("cannot open ") + file; return PGS_ERR_STATE; }
  std::string text; char buf[65536]; size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
  fclose(f);
  pgs::Json obj; std::string perr;
  if (!pgs::Json::parse(text, &obj, &perr)) { h->err = perr; return PGS_ERR_STATE; }
  return h->manager.getWorldsPtr()->loadStateFromDisk(obj.at("WorldsData"), &h->err) ? PGS_OK : PGS_ERR_STATE;
} CATCH_FACADE(h)
int pgs_facade_load_state_from_disk(pgs_facade_handle h, const char* dir) try {
  if (!h || !dir) return PGS_ERR_INVALID_ARGUMENT;
  h->err.clear();
  if (h->manager.getNodeLen() != 0) { h->err = "load_state_from_disk: the session must be empty"; return PGS_ERR_STATE; }
  const std::string file = std::string(dir) + "/solved_posegraph.json";
  if (int rc = pgs_facade_load_worlds_state(h, file.c_str())) return rc;                                   // Composer.cpp:1137
  pgs::SolvedPoseGraph pg;
  if (!pgs::loadSolvedPoseGraph(file, &pg, &h->err)) return PGS_ERR_STATE;
  if (!h->manager.load_kidnap_data(pg.kidnap_starts, pg.kidnap_ends)) { h->err = "load_state_from_disk: kidnap_starts / kidnap_ends do not belong together"; return PGS_ERR_STATE; }   // :1148
  for (size_t i = 0; i < pg.w_T_c.size(); ++i)                                                            // :1158
    if (!h->manager.load_solved_node(pg.stamp_ns[i], pg.w_T_c[i], pg.world_id[i], pg.set_id[i], &h->err)) { h->err = "SolvedPoseGraph[" + std::to_string(i) + "]: " + h->err; return PGS_ERR_STATE; }
  return h->slam->load_state() ? PGS_OK : PGS_ERR_STATE;                                                  // :1167
} CATCH_FACADE(h)
static int copy_out(const std::string& s, char* out, int32_t cap) {
  if (out && cap > 0) { const size_t n = std::min((size_t)cap - 1, s.size()); std::memcpy(out, s.data(), n); out[n] = 0; }
  return (int)s.size();
}
int pgs_io_prettyprint(const double* T16, char* out, int32_t cap) try {
  if (!T16) return PGS_ERR_INVALID_ARGUMENT;
  pgs::Matrix4d T; std::memcpy(T.m, T16, 128);
  return copy_out(pgs::prettyprintMatrix4d(T), out, cap);
} CATCH_IO
int pgs_io_mat_to_string(const double* T16, int32_t solved_layout, char* out, int32_t cap) try {
  if (!T16) return PGS_ERR_INVALID_ARGUMENT;
  pgs::Matrix4d T; std::memcpy(T.m, T16, 128);
  return copy_out(solved_layout ? pgs::mat_to_string(T, ", ", "\n") : pgs::mat_to_string(T), out, cap);
} CATCH_IO
int pgs_io_string_to_mat(const char* s, double* T16) try {
  if (!s || !T16) return PGS_ERR_INVALID_ARGUMENT;
  pgs::Matrix4d T;
  if (!pgs::string_to_mat(s, T)) return 0;
  std::memcpy(T16, T.m, 128);
  return 1;
} CATCH_IO
int pgs_io_load_solved_posegraph(const char* file, double* T, int64_t* stamp_ns, int32_t* world_id, int32_t* set_id, int32_t cap) try {
  if (!file) return PGS_ERR_INVALID_ARGUMENT;
  pgs::SolvedPoseGraph g; std::string err;
  if (!pgs::loadSolvedPoseGraph(file, &g, &err)) return PGS_ERR_STATE;
  const int n = (int)g.w_T_c.size();
  for (int i = 0; i < n && i < cap; ++i) {
    if (T) std::memcpy(T + 16 * (size_t)i, g.w_T_c[i].m, 128);
    if (stamp_ns) stamp_ns[i] = g.stamp_ns[i];
    if (world_id) world_id[i] = g.world_id[i];
    if (set_id) set_id[i] = g.set_id[i];
  }
  return n;
} CATCH_IO
// the term lists are written by the solver thread without a lock of their own: introspection only while it is stopped
static bool introspection_blocked(pgs_facade_handle h) {
  if (!h->solver_thread.joinable()) return false;
  h->err = "introspection of the residual-block lists is not available while the solver thread runs (pgs_facade_thread_stop first)";
  return true;
}
int pgs_facade_alternative_terms_size(pgs_facade_handle h, int32_t kind, int32_t* n_nodes, int32_t* n_edges) try {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  if (introspection_blocked(h)) return PGS_ERR_STATE;
  pgs::PoseGraphSLAM::AlternativeTerms A;
  if (!h->slam->alternative_terms(kind, A)) { h->err = "alternative terms: kind must be 0, 1 or 2"; return PGS_ERR_INVALID_ARGUMENT; }
  if (n_nodes) *n_nodes = A.n_nodes;
  if (n_edges) *n_edges = (int32_t)A.c1.size();
  return PGS_OK;
} CATCH_FACADE(h)
int pgs_facade_get_alternative_terms(pgs_facade_handle h, int32_t kind, int32_t cap_nodes, int32_t cap_edges, double* rot, double* t, int32_t* c1, int32_t* c2,
                                     double* obs_rot, double* obs_t, double* weight, double* sw) try {
  if (!h || cap_nodes < 0 || cap_edges < 0) return PGS_ERR_INVALID_ARGUMENT;
  if (introspection_blocked(h)) return PGS_ERR_STATE;
  pgs::PoseGraphSLAM::AlternativeTerms A;
  if (!h->slam->alternative_terms(kind, A)) { h->err = "alternative terms: kind must be 0, 1 or 2"; return PGS_ERR_INVALID_ARGUMENT; }
  const size_t nn = (size_t)std::min<int>(A.n_nodes, cap_nodes), ne = std::min<size_t>(A.c1.size(), (size_t)cap_edges), rw = kind == 2 ? 3 : 4;
  auto put = [](auto* dst, const auto& v, size_t count) { count = std::min(count, v.size()); if (dst && count) std::memcpy(dst, v.data(), sizeof(v[0]) * count); };
  put(rot, A.rot, rw * nn); put(t, A.t, 3 * nn); put(c1, A.c1, ne); put(c2, A.c2, ne); put(obs_rot, A.obs_rot, rw * ne); put(obs_t, A.obs_t, 3 * ne);
  put(weight, A.weight, ne); put(sw, A.sw, ne);
  return (int)ne;
} CATCH_FACADE(h)
int pgs_facade_evaluate_alternative(pgs_facade_handle h, int32_t kind, int32_t cap_edges, double* r, double* J, double* cost) try {
  if (!h) return PGS_ERR_INVALID_ARGUMENT;
  if (introspection_blocked(h)) return PGS_ERR_STATE;
  pgs::PoseGraphSLAM::AlternativeTerms A;
  if (!h->slam->alternative_terms(kind, A)) { h->err = "alternative terms: kind must be 0, 1 or 2"; return PGS_ERR_INVALID_ARGUMENT; }
  if ((int64_t)A.c1.size() > (int64_t)cap_edges) { h->err = "evaluate_alternative: " + std::to_string(A.c1.size()) + " blocks, room for " + std::to_string(cap_edges); return PGS_ERR_INVALID_ARGUMENT; }
  pgs_fourdof_handle f = nullptr;
  if (int rc = pgs_fourdof_create(h->device, &f)) { h->err = "pgs_fourdof_create: no usable CUDA device (there is no CPU fallback)"; return rc; }
  pgs_fourdof_input in{};
  in.kind = kind; in.n_nodes = A.n_nodes; in.rot = A.rot.data(); in.t = A.t.data(); in.n_edges = (int32_t)A.c1.size(); in.c1 = A.c1.data(); in.c2 = A.c2.data();
  in.obs_rot = A.obs_rot.data(); in.obs_t = A.obs_t.data(); in.weight = A.weight.data(); in.sw = A.sw.data();
  const int rc = pgs_fourdof_evaluate(f, &in, r, J, cost);
  if (rc) h->err = pgs_fourdof_last_error(f);
  pgs_fourdof_destroy(f);
  return rc;
} CATCH_FACADE(h)
int32_t pgs_facade_n_odom_terms(pgs_facade_handle h) { if (!h) return 0; if (introspection_blocked(h)) return PGS_ERR_STATE; return (int32_t)h->slam->odometry_terms().size(); }
int pgs_facade_get_odom_terms(pgs_facade_handle h, int32_t cap, int32_t* u, int32_t* umf, double* q, double* t, double* w) try {
  if (!h || cap < 0) return PGS_ERR_INVALID_ARGUMENT;
  if (introspection_blocked(h)) return PGS_ERR_STATE;
  const auto& v = h->slam->odometry_terms();
  const size_t n = std::min(v.size(), (size_t)cap);
  for (size_t i = 0; i < n; ++i) {
    if (u) u[i] = v[i].u;
    if (umf) umf[i] = v[i].umf;
    if (w) w[i] = v[i].weight;
    if (q) std::memcpy(q + 4 * i, v[i].q, 32);
    if (t) std::memcpy(t + 3 * i, v[i].t, 24);
  }
  return (int)n;
} CATCH_FACADE(h)
int32_t pgs_facade_n_reg_terms(pgs_facade_handle h) { if (!h) return 0; if (introspection_blocked(h)) return PGS_ERR_STATE; return (int32_t)h->slam->regularization_terms().size(); }
int pgs_facade_get_reg_terms(pgs_facade_handle h, int32_t cap, int32_t* node, double* q, double* t, double* w) try {
  if (!h || cap < 0) return PGS_ERR_INVALID_ARGUMENT;
  if (introspection_blocked(h)) return PGS_ERR_STATE;
  const auto& v = h->slam->regularization_terms();
  const size_t n = std::min(v.size(), (size_t)cap);
  for (size_t i = 0; i < n; ++i) {
    double qq[4], tt[3]; pgs::mat_to_raw_xyzw(v[i].anchor, qq, tt);
    if (node) node[i] = v[i].node;
    if (w) w[i] = v[i].weight;
    if (q) std::memcpy(q + 4 * i, qq, 32);
    if (t) std::memcpy(t + 3 * i, tt, 24);
  }
  return (int)n;
} CATCH_FACADE(h)
int32_t pgs_facade_which_world(pgs_facade_handle h, int64_t stamp) { return h ? h->manager.which_world_is_this(stamp) : PGS_ERR_INVALID_ARGUMENT; }
int32_t pgs_facade_n_worlds(pgs_facade_handle h) { return h ? h->manager.n_worlds() : 0; }
int32_t pgs_facade_world_setid(pgs_facade_handle h, int32_t w) { return h ? h->manager.getWorldsPtr()->find_setID_of_world_i(w) : -1; }
int32_t pgs_facade_world_start(pgs_facade_handle h, int32_t w) { return h ? h->manager.nodeidx_of_world_i_started(w) : -1; }
int32_t pgs_facade_world_end(pgs_facade_handle h, int32_t w) { return h ? h->manager.nodeidx_of_world_i_ended(w) : -1; }
int pgs_facade_pose_between_worlds(pgs_facade_handle h, int32_t m, int32_t n, double* M16) try {
  bool ok = true;
  const pgs::Matrix4d T = h->manager.getWorldsPtr()->getPoseBetweenWorlds(m, n, &ok);
  if (M16) std::memcpy(M16, T.m, 128);
  return ok ? 1 : 0;
} CATCH_FACADE(h)

}  // extern "C"

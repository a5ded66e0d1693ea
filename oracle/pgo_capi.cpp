// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see pgo_core.hpp).
// extern "C" surface of the CPU oracle so tests/ and bench.py (cpu_baseline / reference arm)
// can drive it through ctypes.  Nothing under solve_keyframe_pose_graph_b200/ links this.
#include <chrono>
#include <cstdlib>
#include "pgo_solver.hpp"
#include "pgo_fourdof.hpp"

namespace pgo {
double wall_seconds() {
  using namespace std::chrono;
  return duration_cast<duration<double>>(steady_clock::now().time_since_epoch()).count();
}
}  // namespace pgo

using namespace pgo;

extern "C" {

struct pgo_options {
  int max_num_iterations;
  double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius;
  double min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
  int max_num_consecutive_invalid_steps;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  int jacobi_scaling, use_autodiff, num_threads;
};
struct pgo_summary {
  double initial_cost, final_cost;
  int termination, num_successful_steps, num_unsuccessful_steps, num_iterations;
  double t_evaluate, t_linear, t_total;
  double fixed_cost;
};
struct pgo_iteration {  // one row of Ceres' minimizer_progress_to_stdout table
  int iteration; double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, relative_decrease, trust_region_radius;
  int step_is_valid, step_is_successful;
};

void pgo_default_options(pgo_options* o) {
  Options d;
  o->max_num_iterations = d.max_num_iterations; o->initial_trust_region_radius = d.initial_trust_region_radius;
  o->max_trust_region_radius = d.max_trust_region_radius; o->min_trust_region_radius = d.min_trust_region_radius;
  o->min_relative_decrease = d.min_relative_decrease; o->min_lm_diagonal = d.min_lm_diagonal; o->max_lm_diagonal = d.max_lm_diagonal;
  o->max_num_consecutive_invalid_steps = d.max_num_consecutive_invalid_steps; o->function_tolerance = d.function_tolerance;
  o->gradient_tolerance = d.gradient_tolerance; o->parameter_tolerance = d.parameter_tolerance; o->jacobi_scaling = d.jacobi_scaling;
  o->use_autodiff = d.use_autodiff; o->num_threads = d.num_threads;
}
static Options to_opt(const pgo_options* o) {
  Options d;
  if (!o) return d;
  d.max_num_iterations = o->max_num_iterations; d.initial_trust_region_radius = o->initial_trust_region_radius;
  d.max_trust_region_radius = o->max_trust_region_radius; d.min_trust_region_radius = o->min_trust_region_radius;
  d.min_relative_decrease = o->min_relative_decrease; d.min_lm_diagonal = o->min_lm_diagonal; d.max_lm_diagonal = o->max_lm_diagonal;
  d.max_num_consecutive_invalid_steps = o->max_num_consecutive_invalid_steps; d.function_tolerance = o->function_tolerance;
  d.gradient_tolerance = o->gradient_tolerance; d.parameter_tolerance = o->parameter_tolerance; d.jacobi_scaling = o->jacobi_scaling;
  d.use_autodiff = o->use_autodiff; d.num_threads = o->num_threads > 0 ? o->num_threads : 1;
  return d;
}

void* pgo_create() { return new Problem(); }
void pgo_destroy(void* h) { delete (Problem*)h; }

void pgo_set_nodes(void* h, int n, const double* q, const double* t) {
  Problem& P = *(Problem*)h; P.N = n; P.q.assign(q, q + 4 * (size_t)n); P.t.assign(t, t + 3 * (size_t)n); P.node_const.clear();
}
void pgo_set_constant_nodes(void* h, int first, int n, int constant) {
  Problem& P = *(Problem*)h;
  if ((int)P.node_const.size() < P.N) P.node_const.resize(P.N, 0);
  for (int i = first; i < first + n && i < P.N; ++i) P.node_const[i] = constant ? 1 : 0;
}
void pgo_add_odom_edges(void* h, int m, const int* c1, const int* c2, const double* q, const double* t, const double* w) {
  Problem& P = *(Problem*)h;
  P.oc1.insert(P.oc1.end(), c1, c1 + m); P.oc2.insert(P.oc2.end(), c2, c2 + m);
  P.oq.insert(P.oq.end(), q, q + 4 * (size_t)m); P.ot.insert(P.ot.end(), t, t + 3 * (size_t)m); P.ow.insert(P.ow.end(), w, w + m);
}
// Parameters are bound as (c1, c2, switch).  The reference binds a loop edge (a,b) with
// observation b_T_a as (c1,c2) = (b,a)  (src/PoseGraphSLAM.cpp:1550-1556); callers do that swap.
// Each edge gets a fresh switch initialised to s_init[i] (0.99 if null, src/PoseGraphSLAM.cpp:353).
void pgo_add_loop_edges(void* h, int m, const int* c1, const int* c2, const double* q, const double* t, const double* w, const double* s_init) {
  Problem& P = *(Problem*)h;
  for (int i = 0; i < m; ++i) { P.lsi.push_back((int)P.sw.size()); P.sw.push_back(s_init ? s_init[i] : 0.99); }
  P.lc1.insert(P.lc1.end(), c1, c1 + m); P.lc2.insert(P.lc2.end(), c2, c2 + m);
  P.lq.insert(P.lq.end(), q, q + 4 * (size_t)m); P.lt.insert(P.lt.end(), t, t + 3 * (size_t)m); P.lw.insert(P.lw.end(), w, w + m);
}
void pgo_set_regularizers(void* h, int k, const int* node, const double* q, const double* t, const double* w) {
  Problem& P = *(Problem*)h;
  P.rn.assign(node, node + k); P.rq.assign(q, q + 4 * (size_t)k); P.rt.assign(t, t + 3 * (size_t)k); P.rw.assign(w, w + k);
}
void pgo_set_switches(void* h, int m, const double* s) { Problem& P = *(Problem*)h; P.sw.assign(s, s + m); }
void pgo_get_poses(void* h, double* q, double* t) {
  Problem& P = *(Problem*)h; std::memcpy(q, P.q.data(), sizeof(double) * P.q.size()); std::memcpy(t, P.t.data(), sizeof(double) * P.t.size());
}
void pgo_get_switches(void* h, double* s) { Problem& P = *(Problem*)h; std::memcpy(s, P.sw.data(), sizeof(double) * P.sw.size()); }

// Ceres Problem::Evaluate semantics at the current parameters: cost, residual blocks, tangent
// Jacobian blocks (row-major 6x12 / 7x13 / 6x6) and gradient.  Any output may be null.
double pgo_evaluate(void* h, int use_autodiff, int num_threads, double* r_o, double* J_o, double* r_l, double* J_l,
                    double* r_r, double* J_r, double* grad_p, double* grad_s) {
  Problem& P = *(Problem*)h; Options o; o.use_autodiff = use_autodiff; o.num_threads = num_threads > 0 ? num_threads : 1;
  const bool wj = J_o || J_l || J_r || grad_p || grad_s;
  const double cost = P.evaluate(P.q.data(), P.t.data(), P.sw.data(), wj, o);
  auto cp = [](double* dst, const std::vector<double>& src) { if (dst && !src.empty()) std::memcpy(dst, src.data(), sizeof(double) * src.size()); };
  cp(r_o, P.r_o); cp(r_l, P.r_l); cp(r_r, P.r_r);
  if (wj) { cp(J_o, P.J_o); cp(J_l, P.J_l); cp(J_r, P.J_r); }
  if (grad_p || grad_s) { Solver S(P, o); S.compute_gradient(); cp(grad_p, S.grad_p); cp(grad_s, S.grad_s); }
  return cost;
}

// Timed sweep for bench.py: `reps` evaluations of residuals + Jacobians; returns seconds per sweep (best).
double pgo_time_sweep(void* h, int use_autodiff, int num_threads, int reps, int want_jac) {
  Problem& P = *(Problem*)h; Options o; o.use_autodiff = use_autodiff; o.num_threads = num_threads > 0 ? num_threads : 1;
  double best = 1e300;
  for (int i = 0; i < reps; ++i) { const double t0 = wall_seconds(); P.evaluate(P.q.data(), P.t.data(), P.sw.data(), want_jac != 0, o);
    best = std::min(best, wall_seconds() - t0); }
  return best;
}

// One LM linear step at the current parameters with the given radius (first-iteration scaling):
// returns the *unscaled* tangent step delta (6N poses, nsw switches) and the model cost change.
int pgo_linear_step(void* h, const pgo_options* po, double radius, double* delta_p, double* delta_s, double* model_cost_change) {
  Problem& P = *(Problem*)h; Options o = to_opt(po); Solver S(P, o);
  P.evaluate(P.q.data(), P.t.data(), P.sw.data(), true, o);
  S.squared_column_norms(S.scale_p, S.scale_s);
  for (auto& v : S.scale_p) v = o.jacobi_scaling ? 1.0 / (1.0 + std::sqrt(v)) : 1.0;
  for (auto& v : S.scale_s) v = o.jacobi_scaling ? 1.0 / (1.0 + std::sqrt(v)) : 1.0;
  S.squared_column_norms(S.diag_p, S.diag_s);
  for (size_t i = 0; i < S.diag_p.size(); ++i) S.diag_p[i] = std::min(std::max(S.diag_p[i] * S.scale_p[i] * S.scale_p[i], o.min_lm_diagonal), o.max_lm_diagonal);
  for (size_t i = 0; i < S.diag_s.size(); ++i) S.diag_s[i] = std::min(std::max(S.diag_s[i] * S.scale_s[i] * S.scale_s[i], o.min_lm_diagonal), o.max_lm_diagonal);
  std::vector<double> rhs;
  if (!S.compute_step(radius, rhs)) return 1;
  if (model_cost_change) *model_cost_change = S.model_cost_change();
  for (size_t i = 0; i < S.step_p.size(); ++i) delta_p[i] = S.step_p[i] * S.scale_p[i];
  for (size_t i = 0; i < S.step_s.size(); ++i) delta_s[i] = S.step_s[i] * S.scale_s[i];
  return 0;
}

int pgo_solve(void* h, const pgo_options* po, pgo_summary* out, pgo_iteration* iters, int iters_cap) {
  Problem& P = *(Problem*)h; Options o = to_opt(po); Solver S(P, o);
  Summary s = S.solve();
  if (out) { out->initial_cost = s.initial_cost; out->final_cost = s.final_cost; out->termination = s.termination;
    out->num_successful_steps = s.num_successful_steps; out->num_unsuccessful_steps = s.num_unsuccessful_steps;
    out->num_iterations = (int)s.iterations.size(); out->t_evaluate = s.t_evaluate; out->t_linear = s.t_linear; out->t_total = s.t_total; out->fixed_cost = s.fixed_cost; }
  for (int i = 0; iters && i < (int)s.iterations.size() && i < iters_cap; ++i) {
    const IterRecord& r = s.iterations[i];
    iters[i] = pgo_iteration{r.iteration, r.cost, r.cost_change, r.gradient_max_norm, r.gradient_norm, r.step_norm, r.relative_decrease,
                             r.trust_region_radius, r.step_is_valid, r.step_is_successful};
  }
  return s.termination;
}

// ---- small helpers exposed for the Python front-end restatement and the KATs ----
void pgo_mat4_to_pose(const double* M16, double* q, double* t) { Mat4<double> M; std::memcpy(&M.m[0][0], M16, 128); mat4_to_pose(M, q, t); }
void pgo_pose_to_mat4(const double* q, const double* t, double* M16) { Mat4<double> M = pose_to_mat4(q, t); std::memcpy(M16, &M.m[0][0], 128); }
void pgo_inv4(const double* M16, double* out16) { Mat4<double> M; std::memcpy(&M.m[0][0], M16, 128); Mat4<double> I = inv4(M); std::memcpy(out16, &I.m[0][0], 128); }
void pgo_r2ypr_deg(const double* M16, double* ypr) { Mat4<double> M; std::memcpy(&M.m[0][0], M16, 128); R2ypr_deg(M, ypr); }
void pgo_quat_plus(const double* x, const double* d, double* xp) { quat_plus(x, d, xp); }
void pgo_quat_plus_jacobian(const double* x, double* J12) { double J[4][3]; quat_plus_jacobian(x, J); std::memcpy(J12, J, 96); }
// single-block evaluations (mode: 1 autodiff, 0 closed form)
void pgo_sixdof(int mode, const double* q1, const double* t1, const double* q2, const double* t2, const double* oq, const double* ot,
                double w, double* r, double* J) {
  SixDOFError f{Quat<double>{oq[0], oq[1], oq[2], oq[3]}, Vec3<double>{ot[0], ot[1], ot[2]}, w};
  if (mode) eval_sixdof_autodiff(f, q1, t1, q2, t2, r, J); else eval_sixdof_closed(f, q1, t1, q2, t2, r, J);
}
void pgo_sixdof_switch(int mode, const double* q1, const double* t1, const double* q2, const double* t2, const double* s, const double* oq,
                       const double* ot, double w, double* r, double* J) {
  SixDOFErrorWithSwitchingConstraints f{Quat<double>{oq[0], oq[1], oq[2], oq[3]}, Vec3<double>{ot[0], ot[1], ot[2]}, w};
  if (mode) eval_switch_autodiff(f, q1, t1, q2, t2, s, r, J); else eval_switch_closed(f, q1, t1, q2, t2, s, r, J);
}
void pgo_node_reg(int mode, const double* q1, const double* t1, const double* qf, const double* tf, double w, double* r, double* J) {
  NodePoseRegularization f{pose_to_mat4(qf, tf), w};
  if (mode) eval_reg_autodiff(f, q1, t1, r, J); else eval_reg_closed(f, q1, t1, r, J);
}
// ---- the reference's alternative (switched-off) functors, batched over an edge list (pgo_fourdof.hpp).
// kind 0 FourDOFError, 1 FourDOFErrorWithSwitchingConstraints, 2 QinFourDOFWeightError; array meanings as in
// include/pgs_fourdof.h.  J may be null.  Returns 1/2 sum r^2.
double pgo_fourdof_eval(int kind, int n_nodes, const double* rot, const double* t, int n_edges, const int* c1, const int* c2, const double* obs_rot,
                        const double* obs_t, const double* weight, const double* sw, double* r, double* J) {
  (void)n_nodes;
  const int NR = kind == 0 ? 6 : kind == 1 ? 7 : 4, NC = kind == 0 ? 12 : kind == 1 ? 13 : 8;
  double cost = 0;
  for (int e = 0; e < n_edges; ++e) {
    const int a = c1[e], b = c2[e];
    double* re = r + (size_t)NR * e; double* Je = J ? J + (size_t)NR * NC * e : nullptr;
    if (kind == 2) {
      QinFourDOFWeightError f{obs_t[3 * e], obs_t[3 * e + 1], obs_t[3 * e + 2], obs_rot[3 * e], obs_rot[3 * e + 1], obs_rot[3 * e + 2]};
      eval_qin_autodiff(f, rot + 3 * (size_t)a, t + 3 * (size_t)a, rot + 3 * (size_t)b, t + 3 * (size_t)b, re, Je);
    } else {
      const Quat<double> oq{obs_rot[4 * e], obs_rot[4 * e + 1], obs_rot[4 * e + 2], obs_rot[4 * e + 3]};
      const Vec3<double> ot{obs_t[3 * e], obs_t[3 * e + 1], obs_t[3 * e + 2]};
      if (kind == 0) { FourDOFError f{oq, ot, weight[e]}; eval_fourdof_autodiff(f, rot + 4 * (size_t)a, t + 3 * (size_t)a, rot + 4 * (size_t)b, t + 3 * (size_t)b, re, Je); }
      else { FourDOFErrorWithSwitchingConstraints f{oq, ot, weight ? weight[e] : 1.0};
        eval_fourdof_switch_autodiff(f, rot + 4 * (size_t)a, t + 3 * (size_t)a, rot + 4 * (size_t)b, t + 3 * (size_t)b, sw + e, re, Je); }
    }
    for (int i = 0; i < NR; ++i) cost += re[i] * re[i];
  }
  return 0.5 * cost;
}
double pgo_angle_plus(double theta, double delta) { return angle_plus(theta, delta); }
double pgo_angle_plus_jacobian(double theta) { return angle_plus_jacobian(theta); }
// rawyprt_to_eigenmat's rotation (YawPitchRollToRotationMatrix, degrees) for the KATs: R[9] row-major
void pgo_ypr_to_R(double yaw, double pitch, double roll, double* R9) { YawPitchRollToRotationMatrix(yaw, pitch, roll, R9); }

int pgo_max_threads() {
  const unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

}  // extern "C"

// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see pgo_core.hpp).
//
// CPU restatement of what happens inside `ceres::Solve(reint_options, &reint_problem, ...)`
// at src/PoseGraphSLAM.cpp:1903 with the options of src/PoseGraphSLAM.cpp:1268-1272
// (SPARSE_NORMAL_CHOLESKY, max_num_iterations = 10, everything else Ceres defaults).
// Ceres itself is not in the reference tree; this follows Ceres 1.12-1.14
// [CERES-UPSTREAM]: trust_region_minimizer.cc (Minimize / IterationZero /
// ComputeTrustRegionStep / HandleInvalidStep / ParameterToleranceReached /
// FunctionToleranceReached / IsStepSuccessful / HandleSuccessfulStep /
// HandleUnsuccessfulStep / FinalizeIterationAndCheckIfMinimizerCanContinue),
// levenberg_marquardt_strategy.cc (ComputeStep / StepAccepted / StepRejected /
// StepIsInvalid), trust_region_step_evaluator.cc (monotonic case) and
// sparse_normal_cholesky_solver.cc ((J^T J + D^T D) y = J^T r, step = -y).
//
// The linear solve is an exact sparse Cholesky: the scalar switch unknowns are pivoted
// first (their Schur complement adds no fill), then the 6N pose unknowns are factored
// with a skyline (row-envelope) LL^T in node order.  Any exact ordering gives Ceres'
// step up to rounding.
#pragma once
#include <algorithm>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>
#include <thread>
#include "pgo_core.hpp"

namespace pgo {

struct Options {  // ceres::Solver::Options in effect
  int max_num_iterations = 10;               // PoseGraphSLAM.cpp:1272
  double initial_trust_region_radius = 1e4;
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double min_lm_diagonal = 1e-6;
  double max_lm_diagonal = 1e32;
  int max_num_consecutive_invalid_steps = 5;
  double function_tolerance = 1e-6;
  double gradient_tolerance = 1e-10;
  double parameter_tolerance = 1e-8;
  int jacobi_scaling = 1;
  int use_autodiff = 1;   // 1: Jet autodiff (what the reference runs); 0: closed-form Jacobians
  int num_threads = 1;    // Ceres default; the reference never sets it
};

enum Termination { CONVERGENCE = 0, NO_CONVERGENCE = 1, FAILURE = 2 };

struct IterRecord {
  int iteration; double cost, cost_change, gradient_max_norm, gradient_norm, step_norm, relative_decrease,
      trust_region_radius; int step_is_valid, step_is_successful;
};

struct Summary {
  double initial_cost = 0, final_cost = 0;
  double fixed_cost = 0;   // ceres::Solver::Summary::fixed_cost: residual blocks whose parameter blocks are all constant
  int termination = NO_CONVERGENCE;
  int num_successful_steps = 0, num_unsuccessful_steps = 0;
  std::string message;
  std::vector<IterRecord> iterations;
  double t_evaluate = 0, t_linear = 0, t_total = 0;
};

// Static-chunk parallel loop over [0,n) on `nt` std::threads (no OpenMP runtime needed);
// body(begin, end) returns a partial sum, partials are added in thread order (deterministic).
template <class F> inline double parallel_sum(int n, int nt, F body) {
  if (nt <= 1 || n < 4 * nt) return body(0, n);
  std::vector<double> part(nt, 0.0); std::vector<std::thread> th;
  for (int k = 0; k < nt; ++k) { const int b = (int)((long long)n * k / nt), e = (int)((long long)n * (k + 1) / nt);
    th.emplace_back([&, k, b, e] { part[k] = body(b, e); }); }
  for (auto& t : th) t.join();
  double s = 0; for (double p : part) s += p; return s;
}

struct Problem {
  int N = 0;
  std::vector<double> q, t;                 // 4N, 3N  (x,y,z,w)
  // SixDOFError blocks, parameters (c1, c2)
  std::vector<int> oc1, oc2; std::vector<double> oq, ot, ow;
  // SixDOFErrorWithSwitchingConstraints blocks, parameters (c1, c2, s[sidx])
  std::vector<int> lc1, lc2, lsi; std::vector<double> lq, lt, lw;
  std::vector<double> sw;                   // switch variables (one per loop-edge slot)
  // NodePoseRegularization blocks
  std::vector<int> rn; std::vector<double> rq, rt, rw;
  // ceres::Problem::SetParameterBlockConstant on a node's q and t blocks (reference PoseGraphSLAM.cpp:150-151, load_state):
  // Ceres removes constant blocks from the reduced program, i.e. their Jacobian columns vanish and they leave the
  // state vector whose norms drive the convergence tests.  Empty = no constant block.
  std::vector<char> node_const;
  bool is_const(int i) const { return i < (int)node_const.size() && node_const[i]; }

  int n_odom() const { return (int)oc1.size(); }
  int n_loop() const { return (int)lc1.size(); }
  int n_reg() const { return (int)rn.size(); }

  // scratch filled by evaluate()
  std::vector<double> r_o, J_o, r_l, J_l, r_r, J_r;  // 6,72 | 7,91 | 6,36 per block

  // cost = 1/2 sum r^2.  want_jac: fill J_* too.  keep_residuals = false: cost only, r_* stay as they are —
  // Ceres evaluates the candidate point of an LM step with residuals == NULL (TrustRegionMinimizer), so the residuals of
  // the current point x survive a rejected step and the next step is built from J(x)^T r(x).
  double evaluate(const double* Q, const double* T, const double* S, bool want_jac, const Options& opt, bool keep_residuals = true) {
    const int Eo = n_odom(), El = n_loop(), K = n_reg();
    if (keep_residuals) { r_o.resize(6 * (size_t)Eo); r_l.resize(7 * (size_t)El); r_r.resize(6 * (size_t)K); }
    if (want_jac) { J_o.resize(72 * (size_t)Eo); J_l.resize(91 * (size_t)El); J_r.resize(36 * (size_t)K); }
    double cost = 0.0;
    const bool ad = opt.use_autodiff != 0;
    cost += parallel_sum(Eo, opt.num_threads, [&](int e0, int e1) { double cost = 0;
    for (int e = e0; e < e1; ++e) {
      SixDOFError f{Quat<double>{oq[4 * e], oq[4 * e + 1], oq[4 * e + 2], oq[4 * e + 3]},
                    Vec3<double>{ot[3 * e], ot[3 * e + 1], ot[3 * e + 2]}, ow[e]};
      double rtmp[6];
      double* r = keep_residuals ? &r_o[6 * (size_t)e] : rtmp;
      double* J = want_jac ? &J_o[72 * (size_t)e] : nullptr;
      const int a = oc1[e], b = oc2[e];
      if (ad) eval_sixdof_autodiff(f, Q + 4 * a, T + 3 * a, Q + 4 * b, T + 3 * b, r, J);
      else eval_sixdof_closed(f, Q + 4 * a, T + 3 * a, Q + 4 * b, T + 3 * b, r, J);
      double s = 0; for (int i = 0; i < 6; ++i) s += r[i] * r[i];
      cost += s;
    }
    return cost; });
    cost += parallel_sum(El, opt.num_threads, [&](int e0, int e1) { double cost = 0;
    for (int e = e0; e < e1; ++e) {
      SixDOFErrorWithSwitchingConstraints f{Quat<double>{lq[4 * e], lq[4 * e + 1], lq[4 * e + 2], lq[4 * e + 3]},
                                            Vec3<double>{lt[3 * e], lt[3 * e + 1], lt[3 * e + 2]}, lw[e]};
      double rtmp[7];
      double* r = keep_residuals ? &r_l[7 * (size_t)e] : rtmp;
      double* J = want_jac ? &J_l[91 * (size_t)e] : nullptr;
      const int a = lc1[e], b = lc2[e];
      if (ad) eval_switch_autodiff(f, Q + 4 * a, T + 3 * a, Q + 4 * b, T + 3 * b, S + lsi[e], r, J);
      else eval_switch_closed(f, Q + 4 * a, T + 3 * a, Q + 4 * b, T + 3 * b, S + lsi[e], r, J);
      double s = 0; for (int i = 0; i < 7; ++i) s += r[i] * r[i];
      cost += s;
    }
    return cost; });
    for (int k = 0; k < K; ++k) {
      NodePoseRegularization f{pose_to_mat4(&rq[4 * k], &rt[3 * k]), rw[k]};
      double rtmp[6];
      double* r = keep_residuals ? &r_r[6 * (size_t)k] : rtmp;
      double* J = want_jac ? &J_r[36 * (size_t)k] : nullptr;
      const int a = rn[k];
      if (ad) eval_reg_autodiff(f, Q + 4 * a, T + 3 * a, r, J);
      else eval_reg_closed(f, Q + 4 * a, T + 3 * a, r, J);
      for (int i = 0; i < 6; ++i) cost += r[i] * r[i];
    }
    if (want_jac && !node_const.empty()) {      // zero the tangent columns of constant parameter blocks
      for (int e = 0; e < Eo; ++e) for (int side = 0; side < 2; ++side) if (is_const(side ? oc2[e] : oc1[e]))
        for (int i = 0; i < 6; ++i) for (int c = 0; c < 6; ++c) J_o[72 * (size_t)e + 12 * i + 6 * side + c] = 0.0;
      for (int e = 0; e < El; ++e) for (int side = 0; side < 2; ++side) if (is_const(side ? lc2[e] : lc1[e]))
        for (int i = 0; i < 7; ++i) for (int c = 0; c < 6; ++c) J_l[91 * (size_t)e + 13 * i + 6 * side + c] = 0.0;
      for (int k = 0; k < K; ++k) if (is_const(rn[k])) for (int i = 0; i < 36; ++i) J_r[36 * (size_t)k + i] = 0.0;
    }
    return 0.5 * cost;
  }
  // Cost of the residual blocks Ceres' preprocessor removes from the reduced program because every parameter block they
  // bind is constant (Program::RemoveFixedBlocks): odometry blocks between two constant keyframes and regularisers on a
  // constant keyframe — a loop block always keeps its free switch.  From the residuals of the last full evaluation.
  double fixed_cost_of_current_residuals() const {
    if (node_const.empty()) return 0.0;
    double c = 0.0;
    for (int e = 0; e < n_odom(); ++e) if (is_const(oc1[e]) && is_const(oc2[e])) for (int i = 0; i < 6; ++i) c += r_o[6 * (size_t)e + i] * r_o[6 * (size_t)e + i];
    for (int k = 0; k < n_reg(); ++k) if (is_const(rn[k])) for (int i = 0; i < 6; ++i) c += r_r[6 * (size_t)k + i] * r_r[6 * (size_t)k + i];
    return 0.5 * c;
  }
};

// ----------------------------------------------------------------------------------
// Skyline (row-envelope) Cholesky, scalar, node-aligned envelopes.
// ----------------------------------------------------------------------------------
struct Skyline {
  int n = 0;
  std::vector<int> start;        // first stored column of row i
  std::vector<size_t> ptr;       // row i occupies val[ptr[i] .. ptr[i] + (i-start[i]) ] (diagonal last)
  std::vector<double> val;
  void init(int n_, const std::vector<int>& start_) {
    n = n_; start = start_; ptr.resize(n + 1); ptr[0] = 0;
    for (int i = 0; i < n; ++i) ptr[i + 1] = ptr[i] + (size_t)(i - start[i] + 1);
    val.assign(ptr[n], 0.0);
  }
  inline double& at(int i, int j) { return val[ptr[i] + (size_t)(j - start[i])]; }  // j in [start[i], i]
  void zero() { std::fill(val.begin(), val.end(), 0.0); }
  static inline double dot(const double* a, const double* b, int len) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0; int k = 0;
    for (; k + 4 <= len; k += 4) { s0 += a[k] * b[k]; s1 += a[k + 1] * b[k + 1]; s2 += a[k + 2] * b[k + 2]; s3 += a[k + 3] * b[k + 3]; }
    for (; k < len; ++k) s0 += a[k] * b[k];
    return (s0 + s1) + (s2 + s3);
  }
  // In-place A = L L^T.  Returns false on a non-positive pivot.
  bool factor() {
    for (int i = 0; i < n; ++i) {
      double* Li = &val[ptr[i]]; const int si = start[i];
      for (int j = si; j < i; ++j) {
        const int sj = start[j]; const int k0 = std::max(si, sj);
        const double* Lj = &val[ptr[j]];
        const double s = Li[j - si] - dot(Li + (k0 - si), Lj + (k0 - sj), j - k0);
        Li[j - si] = s / Lj[j - sj];
      }
      const double d = Li[i - si] - dot(Li, Li, i - si);
      if (!(d > 0.0) || !std::isfinite(d)) return false;
      Li[i - si] = std::sqrt(d);
    }
    return true;
  }
  void solve(double* b) const {  // in place
    for (int i = 0; i < n; ++i) {
      const double* Li = &val[ptr[i]]; const int si = start[i];
      b[i] = (b[i] - dot(Li, b + si, i - si)) / Li[i - si];
    }
    for (int i = n - 1; i >= 0; --i) {
      const double* Li = &val[ptr[i]]; const int si = start[i];
      const double x = b[i] / Li[i - si]; b[i] = x;
      for (int k = si; k < i; ++k) b[k] -= Li[k - si] * x;
    }
  }
};

double wall_seconds();

// ----------------------------------------------------------------------------------
// Trust-region Levenberg-Marquardt, Ceres 1.12-1.14 semantics.
// ----------------------------------------------------------------------------------
struct Solver {
  Problem& P; Options opt;
  int N, Eo, El, K;
  std::vector<char> node_used, sw_used;
  std::vector<double> scale_p, scale_s;   // Jacobi scaling (6N, nsw)
  std::vector<double> diag_p, diag_s;     // clamped squared column norms of the scaled Jacobian
  std::vector<double> grad_p, grad_s;     // unscaled gradient J^T r
  Skyline A;
  std::vector<double> step_p, step_s;     // scaled trust-region step
  std::vector<double> lv, lhss, lgs;      // per loop block: v (12), h_ss, g_s

  Solver(Problem& p, const Options& o) : P(p), opt(o) {
    N = P.N; Eo = P.n_odom(); El = P.n_loop(); K = P.n_reg();
    node_used.assign(N, 0); sw_used.assign(P.sw.size(), 0);
    std::vector<int> nstart(N);
    for (int i = 0; i < N; ++i) nstart[i] = i;
    auto touch = [&](int a, int b) { node_used[a] = node_used[b] = 1; const int lo = std::min(a, b), hi = std::max(a, b); nstart[hi] = std::min(nstart[hi], lo); };
    for (int e = 0; e < Eo; ++e) touch(P.oc1[e], P.oc2[e]);
    for (int e = 0; e < El; ++e) { touch(P.lc1[e], P.lc2[e]); sw_used[P.lsi[e]] = 1; }
    for (int k = 0; k < K; ++k) node_used[P.rn[k]] = 1;
    for (int i = 0; i < N; ++i) if (P.is_const(i)) node_used[i] = 0;   // constant blocks are not part of the reduced program
    std::vector<int> start(6 * (size_t)N);
    for (int i = 0; i < N; ++i) for (int c = 0; c < 6; ++c) start[6 * i + c] = 6 * nstart[i];
    A.init(6 * N, start);
  }

  void compute_gradient() {
    grad_p.assign(6 * (size_t)N, 0.0); grad_s.assign(P.sw.size(), 0.0);
    for (int e = 0; e < Eo; ++e) {
      const double* J = &P.J_o[72 * (size_t)e]; const double* r = &P.r_o[6 * (size_t)e];
      for (int c = 0; c < 12; ++c) { double s = 0; for (int i = 0; i < 6; ++i) s += J[12 * i + c] * r[i];
        grad_p[6 * (c < 6 ? P.oc1[e] : P.oc2[e]) + c % 6] += s; }
    }
    for (int e = 0; e < El; ++e) {
      const double* J = &P.J_l[91 * (size_t)e]; const double* r = &P.r_l[7 * (size_t)e];
      for (int c = 0; c < 12; ++c) { double s = 0; for (int i = 0; i < 7; ++i) s += J[13 * i + c] * r[i];
        grad_p[6 * (c < 6 ? P.lc1[e] : P.lc2[e]) + c % 6] += s; }
      double s = 0; for (int i = 0; i < 7; ++i) s += J[13 * i + 12] * r[i];
      grad_s[P.lsi[e]] += s;
    }
    for (int k = 0; k < K; ++k) {
      const double* J = &P.J_r[36 * (size_t)k]; const double* r = &P.r_r[6 * (size_t)k];
      for (int c = 0; c < 6; ++c) { double s = 0; for (int i = 0; i < 6; ++i) s += J[6 * i + c] * r[i]; grad_p[6 * P.rn[k] + c] += s; }
    }
  }
  // sum_i J_ij^2 for every tangent column (unscaled J).
  void squared_column_norms(std::vector<double>& cp, std::vector<double>& cs) {
    cp.assign(6 * (size_t)N, 0.0); cs.assign(P.sw.size(), 0.0);
    for (int e = 0; e < Eo; ++e) { const double* J = &P.J_o[72 * (size_t)e];
      for (int c = 0; c < 12; ++c) { double s = 0; for (int i = 0; i < 6; ++i) s += J[12 * i + c] * J[12 * i + c];
        cp[6 * (c < 6 ? P.oc1[e] : P.oc2[e]) + c % 6] += s; } }
    for (int e = 0; e < El; ++e) { const double* J = &P.J_l[91 * (size_t)e];
      for (int c = 0; c < 12; ++c) { double s = 0; for (int i = 0; i < 7; ++i) s += J[13 * i + c] * J[13 * i + c];
        cp[6 * (c < 6 ? P.lc1[e] : P.lc2[e]) + c % 6] += s; }
      double s = 0; for (int i = 0; i < 7; ++i) s += J[13 * i + 12] * J[13 * i + 12]; cs[P.lsi[e]] += s; }
    for (int k = 0; k < K; ++k) { const double* J = &P.J_r[36 * (size_t)k];
      for (int c = 0; c < 6; ++c) { double s = 0; for (int i = 0; i < 6; ++i) s += J[6 * i + c] * J[6 * i + c]; cp[6 * P.rn[k] + c] += s; } }
  }

  // Add the 6x6 block  B (row node a, col node b) into the skyline (lower triangle only).
  inline void add_block(int a, int b, const double* B /*6x6 row-major*/) {
    if (a == b) { for (int i = 0; i < 6; ++i) for (int j = 0; j <= i; ++j) A.at(6 * a + i, 6 * a + j) += B[6 * i + j]; }
    else if (a > b) { for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A.at(6 * a + i, 6 * b + j) += B[6 * i + j]; }
    else { for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A.at(6 * b + j, 6 * a + i) += B[6 * i + j]; }
  }
  // contribution of one residual block with scaled pose Jacobian Js (m x 12) minus rank-1 term v v^T / hss
  void add_pair(int c1, int c2, const double* Js, int m, int ld, const double* v, double inv_hss) {
    double B[36];
    for (int bi = 0; bi < 2; ++bi)
      for (int bj = 0; bj <= bi; ++bj) {
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
          double s = 0; for (int k = 0; k < m; ++k) s += Js[ld * k + 6 * bi + i] * Js[ld * k + 6 * bj + j];
          if (v) s -= v[6 * bi + i] * v[6 * bj + j] * inv_hss;
          B[6 * i + j] = s; }
        add_block(bi ? c2 : c1, bj ? c2 : c1, B);
      }
  }

  // LevenbergMarquardtStrategy::ComputeStep + SparseNormalCholeskySolver.  Returns false on solver failure.
  bool compute_step(double radius, std::vector<double>& rhs_p) {
    A.zero(); rhs_p.assign(6 * (size_t)N, 0.0);
    lv.resize(12 * (size_t)El); lhss.resize(El); lgs.resize(El);
    double Js[7 * 12];
    for (int e = 0; e < Eo; ++e) {
      const double* J = &P.J_o[72 * (size_t)e]; const double* r = &P.r_o[6 * (size_t)e]; const int c1 = P.oc1[e], c2 = P.oc2[e];
      for (int i = 0; i < 6; ++i) for (int c = 0; c < 12; ++c) Js[12 * i + c] = J[12 * i + c] * scale_p[6 * (c < 6 ? c1 : c2) + c % 6];
      add_pair(c1, c2, Js, 6, 12, nullptr, 0.0);
      for (int c = 0; c < 12; ++c) { double s = 0; for (int i = 0; i < 6; ++i) s += Js[12 * i + c] * r[i]; rhs_p[6 * (c < 6 ? c1 : c2) + c % 6] += s; }
    }
    for (int e = 0; e < El; ++e) {
      const double* J = &P.J_l[91 * (size_t)e]; const double* r = &P.r_l[7 * (size_t)e]; const int c1 = P.lc1[e], c2 = P.lc2[e], si = P.lsi[e];
      double js[7];
      for (int i = 0; i < 7; ++i) { for (int c = 0; c < 12; ++c) Js[12 * i + c] = J[13 * i + c] * scale_p[6 * (c < 6 ? c1 : c2) + c % 6];
        js[i] = J[13 * i + 12] * scale_s[si]; }
      double hss = diag_s[si] / radius, gs = 0; double* v = &lv[12 * (size_t)e];
      for (int i = 0; i < 7; ++i) { hss += js[i] * js[i]; gs += js[i] * r[i]; }
      for (int c = 0; c < 12; ++c) { double s = 0; for (int i = 0; i < 7; ++i) s += Js[12 * i + c] * js[i]; v[c] = s; }
      lhss[e] = hss; lgs[e] = gs;
      add_pair(c1, c2, Js, 7, 12, v, 1.0 / hss);
      for (int c = 0; c < 12; ++c) { double s = 0; for (int i = 0; i < 7; ++i) s += Js[12 * i + c] * r[i];
        rhs_p[6 * (c < 6 ? c1 : c2) + c % 6] += s - v[c] * gs / hss; }
    }
    for (int k = 0; k < K; ++k) {
      const double* J = &P.J_r[36 * (size_t)k]; const double* r = &P.r_r[6 * (size_t)k]; const int a = P.rn[k];
      double B[36];
      for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0;
        for (int m = 0; m < 6; ++m) s += J[6 * m + i] * scale_p[6 * a + i] * J[6 * m + j] * scale_p[6 * a + j]; B[6 * i + j] = s; }
      add_block(a, a, B);
      for (int c = 0; c < 6; ++c) { double s = 0; for (int i = 0; i < 6; ++i) s += J[6 * i + c] * scale_p[6 * a + c] * r[i]; rhs_p[6 * a + c] += s; }
    }
    for (int i = 0; i < 6 * N; ++i) {
      if (node_used[i / 6]) A.at(i, i) += diag_p[i] / radius; else A.at(i, i) = 1.0;
    }
    if (!A.factor()) return false;
    step_p = rhs_p; A.solve(step_p.data());
    step_s.assign(P.sw.size(), 0.0);
    for (int e = 0; e < El; ++e) {
      const double* v = &lv[12 * (size_t)e]; const int c1 = P.lc1[e], c2 = P.lc2[e];
      double s = lgs[e];
      for (int c = 0; c < 12; ++c) s -= v[c] * step_p[6 * (c < 6 ? c1 : c2) + c % 6];
      step_s[P.lsi[e]] = s / lhss[e];
    }
    for (auto& x : step_p) { if (!std::isfinite(x)) return false; x = -x; }
    for (auto& x : step_s) { if (!std::isfinite(x)) return false; x = -x; }
    return true;
  }

  // -(J step)^T (r + J step / 2), with the scaled Jacobian and scaled step.
  double model_cost_change() {
    double acc = 0;
    for (int e = 0; e < Eo; ++e) { const double* J = &P.J_o[72 * (size_t)e]; const double* r = &P.r_o[6 * (size_t)e]; const int c1 = P.oc1[e], c2 = P.oc2[e];
      for (int i = 0; i < 6; ++i) { double m = 0; for (int c = 0; c < 12; ++c) { const int u = 6 * (c < 6 ? c1 : c2) + c % 6; m += J[12 * i + c] * scale_p[u] * step_p[u]; }
        acc += m * (r[i] + m / 2.0); } }
    for (int e = 0; e < El; ++e) { const double* J = &P.J_l[91 * (size_t)e]; const double* r = &P.r_l[7 * (size_t)e]; const int c1 = P.lc1[e], c2 = P.lc2[e], si = P.lsi[e];
      for (int i = 0; i < 7; ++i) { double m = J[13 * i + 12] * scale_s[si] * step_s[si];
        for (int c = 0; c < 12; ++c) { const int u = 6 * (c < 6 ? c1 : c2) + c % 6; m += J[13 * i + c] * scale_p[u] * step_p[u]; }
        acc += m * (r[i] + m / 2.0); } }
    for (int k = 0; k < K; ++k) { const double* J = &P.J_r[36 * (size_t)k]; const double* r = &P.r_r[6 * (size_t)k]; const int a = P.rn[k];
      for (int i = 0; i < 6; ++i) { double m = 0; for (int c = 0; c < 6; ++c) m += J[6 * i + c] * scale_p[6 * a + c] * step_p[6 * a + c];
        acc += m * (r[i] + m / 2.0); } }
    return -acc;
  }

  void plus(const std::vector<double>& q, const std::vector<double>& t, const std::vector<double>& s,
            const std::vector<double>& dp, const std::vector<double>& ds, std::vector<double>& q2,
            std::vector<double>& t2, std::vector<double>& s2) {
    q2 = q; t2 = t; s2 = s;
    for (int i = 0; i < N; ++i) {
      if (!node_used[i]) continue;
      quat_plus(&q[4 * i], &dp[6 * i], &q2[4 * i]);
      for (int c = 0; c < 3; ++c) t2[3 * i + c] = t[3 * i + c] + dp[6 * i + 3 + c];
    }
    for (size_t e = 0; e < s.size(); ++e) if (sw_used[e]) s2[e] = s[e] + ds[e];
  }
  double ambient_norm(const std::vector<double>& q, const std::vector<double>& t, const std::vector<double>& s) {
    double a = 0;
    for (int i = 0; i < N; ++i) if (node_used[i]) { for (int c = 0; c < 4; ++c) a += q[4 * i + c] * q[4 * i + c]; for (int c = 0; c < 3; ++c) a += t[3 * i + c] * t[3 * i + c]; }
    for (size_t e = 0; e < s.size(); ++e) if (sw_used[e]) a += s[e] * s[e];
    return std::sqrt(a);
  }
  double ambient_diff_norm(const std::vector<double>& q, const std::vector<double>& t, const std::vector<double>& s,
                           const std::vector<double>& q2, const std::vector<double>& t2, const std::vector<double>& s2, double* maxn) {
    double a = 0, m = 0;
    auto acc = [&](double d) { a += d * d; m = std::max(m, std::fabs(d)); };
    for (int i = 0; i < N; ++i) if (node_used[i]) { for (int c = 0; c < 4; ++c) acc(q[4 * i + c] - q2[4 * i + c]); for (int c = 0; c < 3; ++c) acc(t[3 * i + c] - t2[3 * i + c]); }
    for (size_t e = 0; e < s.size(); ++e) if (sw_used[e]) acc(s[e] - s2[e]);
    if (maxn) *maxn = m;
    return std::sqrt(a);
  }

  Summary solve() {
    Summary sum; const double t0 = wall_seconds();
    std::vector<double> xq = P.q, xt = P.t, xs = P.sw;        // x_
    std::vector<double> cq, ct, cs;                           // candidate_x_
    std::vector<double> rhs_p, dp, ds, nq, nt, ns;
    double radius = opt.initial_trust_region_radius, decrease_factor = 2.0; bool reuse_diagonal = false;
    int num_consecutive_invalid_steps = 0;
    double x_norm = ambient_norm(xq, xt, xs);
    double x_cost = 0, grad_max = 0, grad_norm = 0, fixed_cost = 0;

    auto evaluate_gradient_and_jacobian = [&](int iteration) {
      double te = wall_seconds();
      x_cost = P.evaluate(xq.data(), xt.data(), xs.data(), true, opt);
      if (iteration == 0) fixed_cost = P.fixed_cost_of_current_residuals();   // constant blocks do not move: evaluated once, as Ceres does
      x_cost -= fixed_cost;                                                   // the minimiser sees the reduced program's cost
      compute_gradient();
      if (iteration == 0) {
        if (opt.jacobi_scaling) {
          squared_column_norms(scale_p, scale_s);
          for (auto& v : scale_p) v = 1.0 / (1.0 + std::sqrt(v));
          for (auto& v : scale_s) v = 1.0 / (1.0 + std::sqrt(v));
        } else { scale_p.assign(6 * (size_t)N, 1.0); scale_s.assign(P.sw.size(), 1.0); }
      }
      // |Plus(x, -g) - x|  (projected gradient step, ambient space)
      std::vector<double> ngp(grad_p.size()), ngs(grad_s.size());
      for (size_t i = 0; i < ngp.size(); ++i) ngp[i] = -grad_p[i];
      for (size_t i = 0; i < ngs.size(); ++i) ngs[i] = -grad_s[i];
      plus(xq, xt, xs, ngp, ngs, nq, nt, ns);
      grad_norm = ambient_diff_norm(xq, xt, xs, nq, nt, ns, &grad_max);
      sum.t_evaluate += wall_seconds() - te;
    };

    // ---- IterationZero
    evaluate_gradient_and_jacobian(0);
    sum.initial_cost = x_cost + fixed_cost; sum.fixed_cost = fixed_cost;
    IterRecord it{}; it.iteration = 0; it.cost = x_cost; it.gradient_max_norm = grad_max; it.gradient_norm = grad_norm;
    it.step_is_valid = 1; it.step_is_successful = 1; it.trust_region_radius = radius;

    while (true) {
      // ---- FinalizeIterationAndCheckIfMinimizerCanContinue
      if (it.step_is_successful) ++sum.num_successful_steps; else ++sum.num_unsuccessful_steps;
      it.trust_region_radius = radius;
      sum.iterations.push_back(it);
      if (it.iteration >= opt.max_num_iterations) { sum.termination = NO_CONVERGENCE; sum.message = "Maximum number of iterations reached."; break; }
      if (it.step_is_successful && it.gradient_max_norm <= opt.gradient_tolerance) { sum.termination = CONVERGENCE; sum.message = "Gradient tolerance reached."; break; }
      if (radius <= opt.min_trust_region_radius) { sum.termination = CONVERGENCE; sum.message = "Minimum trust region radius reached."; break; }

      IterRecord prev = it; it = IterRecord{}; it.iteration = prev.iteration + 1;
      it.gradient_max_norm = prev.gradient_max_norm; it.gradient_norm = prev.gradient_norm;

      // ---- ComputeTrustRegionStep
      double tl = wall_seconds();
      if (!reuse_diagonal) {
        squared_column_norms(diag_p, diag_s);
        for (size_t i = 0; i < diag_p.size(); ++i) diag_p[i] = std::min(std::max(diag_p[i] * scale_p[i] * scale_p[i], opt.min_lm_diagonal), opt.max_lm_diagonal);
        for (size_t i = 0; i < diag_s.size(); ++i) diag_s[i] = std::min(std::max(diag_s[i] * scale_s[i] * scale_s[i], opt.min_lm_diagonal), opt.max_lm_diagonal);
      }
      const bool ok = compute_step(radius, rhs_p);
      reuse_diagonal = true;
      sum.t_linear += wall_seconds() - tl;
      double mcc = 0;
      if (ok) { mcc = model_cost_change(); it.step_is_valid = mcc > 0.0; } else it.step_is_valid = 0;
      if (!it.step_is_valid) {
        // ---- HandleInvalidStep
        if (++num_consecutive_invalid_steps >= opt.max_num_consecutive_invalid_steps) {
          sum.termination = FAILURE; sum.message = "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps";
          it.cost = x_cost; it.trust_region_radius = radius; sum.iterations.push_back(it); break; }
        radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;  // StepIsInvalid -> StepRejected(0)
        it.cost = x_cost; it.cost_change = 0; it.step_norm = 0; it.relative_decrease = 0; it.step_is_successful = 0;
        continue;
      }
      num_consecutive_invalid_steps = 0;
      dp.resize(step_p.size()); ds.resize(step_s.size());
      for (size_t i = 0; i < dp.size(); ++i) dp[i] = step_p[i] * scale_p[i];
      for (size_t i = 0; i < ds.size(); ++i) ds[i] = step_s[i] * scale_s[i];

      // ---- ComputeCandidatePointAndEvaluateCost
      plus(xq, xt, xs, dp, ds, cq, ct, cs);
      double te = wall_seconds();
      double candidate_cost = P.evaluate(cq.data(), ct.data(), cs.data(), false, opt, /*keep_residuals=*/false) - fixed_cost;
      sum.t_evaluate += wall_seconds() - te;
      if (!std::isfinite(candidate_cost)) candidate_cost = std::numeric_limits<double>::max();

      // ---- ParameterToleranceReached
      it.step_norm = ambient_diff_norm(xq, xt, xs, cq, ct, cs, nullptr);
      if (it.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
        sum.termination = CONVERGENCE; sum.message = "Parameter tolerance reached."; it.cost = x_cost; sum.iterations.push_back(it); break; }
      // ---- FunctionToleranceReached
      it.cost_change = x_cost - candidate_cost;
      if (std::fabs(it.cost_change) <= opt.function_tolerance * x_cost) {
        sum.termination = CONVERGENCE; sum.message = "Function tolerance reached."; it.cost = x_cost; sum.iterations.push_back(it); break; }
      // ---- IsStepSuccessful
      it.relative_decrease = (x_cost - candidate_cost) / mcc;
      if (it.relative_decrease > opt.min_relative_decrease) {
        // ---- HandleSuccessfulStep
        xq = cq; xt = ct; xs = cs; x_norm = ambient_norm(xq, xt, xs);
        evaluate_gradient_and_jacobian(it.iteration);
        it.cost = x_cost; it.gradient_max_norm = grad_max; it.gradient_norm = grad_norm; it.step_is_successful = 1;
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
        radius = std::min(opt.max_trust_region_radius, radius);
        decrease_factor = 2.0; reuse_diagonal = false;
      } else {
        // ---- HandleUnsuccessfulStep
        it.step_is_successful = 0; it.cost = candidate_cost;
        radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      }
    }
    P.q = xq; P.t = xt; P.sw = xs;
    sum.final_cost = x_cost + fixed_cost;
    sum.t_total = wall_seconds() - t0;
    return sum;
  }
};

}  // namespace pgo

// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// The reference's cost functors, COMPILED FROM THE REFERENCE'S OWN SOURCE FILE where it lies
// (/root/reference/src/CeresResidues.h, unmodified, included below) over oracle/shim/ — a stand-in for the subset of the
// Eigen / Ceres API that file uses, because neither library exists in this container (see oracle/shim/mini_eigen.hpp
// for exactly what that does and does not pin).  oracle/Makefile builds this into oracle/_ref/libref_functors.so;
// tests/test_reference_functors.py compares the oracle's restatement (pgo_core.hpp, pgo_fourdof.hpp) with it.
// The same library holds the reference's src/utils/PoseManipUtils.cpp, compiled from where it lies as a second
// translation unit (conversions between (quaternion, t), (yaw/pitch/roll, t) and 4x4, R2ypr / ypr2R, the
// `data_pretty` printer and the `a,b;c,d` matrix parser), reachable through the ref_pmu_* functions below.
// Differentiation: the functors are instantiated with pgo::Jet (ambient parameters), then every quaternion block is
// multiplied by the 4x3 Plus-Jacobian of EigenQuaternionParameterization — what ceres::AutoDiffCostFunction and the
// local parameterisation do.
#include <cstring>
#include <type_traits>

#include "pgo_core.hpp"           // pgo::Jet with its sqrt / sin / cos / atan2 (found by ADL from the reference's templates), quat_plus_jacobian

#include "CeresResidues.h"        // -I /root/reference/src : the reference's file

namespace {

Matrix4d mat16(const double* M16) { Matrix4d M; for (int i = 0; i < 16; ++i) M.a[i] = M16[i]; return M; }

template <class F, class JT> void call(const F& f, JT* q1, JT* t1, JT* q2, JT* t2, JT*, JT* res, std::false_type) { f(q1, t1, q2, t2, res); }
template <class F, class JT> void call(const F& f, JT* q1, JT* t1, JT* q2, JT* t2, JT* s, JT* res, std::true_type) { f(q1, t1, q2, t2, s, res); }


// residuals and tangent Jacobian of a functor over blocks (q1[4], t1[3], q2[4], t2[3] [, s[1]])
template <int NR, bool SW, class F>
void eval_pair(const F& f, const double* q1, const double* t1, const double* q2, const double* t2, const double* s, double* r, double* J) {
  constexpr int NA = SW ? 15 : 14, NC = SW ? 13 : 12;
  typedef pgo::Jet<NA> JT;
  JT jq1[4], jt1[3], jq2[4], jt2[3], js[1], res[NR];
  for (int i = 0; i < 4; ++i) { jq1[i] = JT(q1[i], i); jq2[i] = JT(q2[i], 7 + i); }
  for (int i = 0; i < 3; ++i) { jt1[i] = JT(t1[i], 4 + i); jt2[i] = JT(t2[i], 11 + i); }
  if (SW) js[0] = JT(s[0], 14);
  call(f, jq1, jt1, jq2, jt2, js, res, std::integral_constant<bool, SW>());
  double P1[4][3], P2[4][3];
  pgo::quat_plus_jacobian(q1, P1); pgo::quat_plus_jacobian(q2, P2);
  for (int i = 0; i < NR; ++i) {
    r[i] = res[i].a;
    if (!J) continue;
    double* Ji = J + NC * i;
    for (int c = 0; c < 3; ++c) {
      Ji[c] = res[i].v[0] * P1[0][c] + res[i].v[1] * P1[1][c] + res[i].v[2] * P1[2][c] + res[i].v[3] * P1[3][c];
      Ji[3 + c] = res[i].v[4 + c];
      Ji[6 + c] = res[i].v[7] * P2[0][c] + res[i].v[8] * P2[1][c] + res[i].v[9] * P2[2][c] + res[i].v[10] * P2[3][c];
      Ji[9 + c] = res[i].v[11 + c];
    }
    if (SW) Ji[12] = res[i].v[14];
  }
}
}  // namespace

extern "C" {

// observations arrive as the row-major 4x4 the reference's constructors take
void ref_sixdof(const double* q1, const double* t1, const double* q2, const double* t2, const double* obs16, double w, double* r, double* J) {
  SixDOFError f(mat16(obs16), w); eval_pair<6, false>(f, q1, t1, q2, t2, nullptr, r, J);
}
void ref_sixdof_switch(const double* q1, const double* t1, const double* q2, const double* t2, const double* s, const double* obs16, double w, double* r, double* J) {
  SixDOFErrorWithSwitchingConstraints f(mat16(obs16), w); eval_pair<7, true>(f, q1, t1, q2, t2, s, r, J);
}
void ref_fourdof(const double* q1, const double* t1, const double* q2, const double* t2, const double* obs16, double w, double* r, double* J) {
  FourDOFError f(mat16(obs16), w); eval_pair<6, false>(f, q1, t1, q2, t2, nullptr, r, J);
}
void ref_fourdof_switch(const double* q1, const double* t1, const double* q2, const double* t2, const double* s, const double* obs16, double w, double* r, double* J) {
  FourDOFErrorWithSwitchingConstraints f(mat16(obs16), w); eval_pair<7, true>(f, q1, t1, q2, t2, s, r, J);
}
void ref_node_reg(const double* q1, const double* t1, const double* pose16, double w, double* r, double* J) {
  NodePoseRegularization f(mat16(pose16), w);
  typedef pgo::Jet<7> JT;
  JT jq[4], jt[3], res[6];
  for (int i = 0; i < 4; ++i) jq[i] = JT(q1[i], i);
  for (int i = 0; i < 3; ++i) jt[i] = JT(t1[i], 4 + i);
  f(jq, jt, res);
  double P[4][3]; pgo::quat_plus_jacobian(q1, P);
  for (int i = 0; i < 6; ++i) {
    r[i] = res[i].a;
    if (!J) continue;
    for (int c = 0; c < 3; ++c) { J[6 * i + c] = res[i].v[0] * P[0][c] + res[i].v[1] * P[1][c] + res[i].v[2] * P[2][c] + res[i].v[3] * P[3][c]; J[6 * i + 3 + c] = res[i].v[4 + c]; }
  }
}
void ref_qin(double yaw_i, const double* ti, double yaw_j, const double* tj, const double* t_obs, double relative_yaw, double pitch_i, double roll_i, double* r, double* J) {
  QinFourDOFWeightError f(t_obs[0], t_obs[1], t_obs[2], relative_yaw, pitch_i, roll_i);
  typedef pgo::Jet<8> JT;
  JT jyi[1], jti[3], jyj[1], jtj[3], res[4];
  jyi[0] = JT(yaw_i, 0); jyj[0] = JT(yaw_j, 4);
  for (int i = 0; i < 3; ++i) { jti[i] = JT(ti[i], 1 + i); jtj[i] = JT(tj[i], 5 + i); }
  f(jyi, jti, jyj, jtj, res);
  // AngleLocalParameterization through autodiff: d Plus / d delta at delta = 0
  AngleLocalParameterization plus;
  pgo::Jet<1> th(yaw_i), d(0.0, 0), out; plus(&th, &d, &out); const double Pi = out.v[0];
  th = pgo::Jet<1>(yaw_j); plus(&th, &d, &out); const double Pj = out.v[0];
  for (int i = 0; i < 4; ++i) {
    r[i] = res[i].a;
    if (!J) continue;
    for (int c = 0; c < 8; ++c) J[8 * i + c] = res[i].v[c] * (c == 0 ? Pi : c == 4 ? Pj : 1.0);
  }
}
// ---- the reference's PoseManipUtils (src/utils/PoseManipUtils.cpp, real source)
void ref_pmu_raw_xyzw_to_eigenmat(const double* q, const double* t, double* M16) { Matrix4d M; PoseManipUtils::raw_xyzw_to_eigenmat(q, t, M); for (int i = 0; i < 16; ++i) M16[i] = M.a[i]; }
void ref_pmu_eigenmat_to_raw_xyzw(const double* M16, double* q, double* t) { PoseManipUtils::eigenmat_to_raw_xyzw(mat16(M16), q, t); }
void ref_pmu_rawyprt_to_eigenmat(const double* ypr, const double* t, double* M16) { Matrix4d M; PoseManipUtils::rawyprt_to_eigenmat(ypr, t, M); for (int i = 0; i < 16; ++i) M16[i] = M.a[i]; }
void ref_pmu_eigenmat_to_rawyprt(const double* M16, double* ypr, double* t) { PoseManipUtils::eigenmat_to_rawyprt(mat16(M16), ypr, t); }
int ref_pmu_prettyprint(const double* M16, char* out, int cap) { const std::string s = PoseManipUtils::prettyprintMatrix4d(mat16(M16)); if ((int)s.size() + 1 > cap) return -1; std::memcpy(out, s.c_str(), s.size() + 1); return (int)s.size(); }
int ref_pmu_string_to_eigenmat(const char* s, double* M16) { Matrix4d M = Matrix4d::Zero(); const bool ok = PoseManipUtils::string_to_eigenmat(std::string(s), M); for (int i = 0; i < 16; ++i) M16[i] = M.a[i]; return ok ? 1 : 0; }

double ref_normalize_angle(double a) { return NormalizeAngle(a); }
double ref_angle_plus(double theta, double delta) { AngleLocalParameterization p; double out; p(&theta, &delta, &out); return out; }
void ref_ypr_to_R(double y, double p, double r, double* R9) { YawPitchRollToRotationMatrix(y, p, r, R9); }
void ref_r2ypr(const double* R9, double* ypr) { Matrix<double, 3, 3> R; for (int i = 0; i < 9; ++i) R.a[i] = R9[i]; Matrix<double, 3, 1> o = R2ypr(R); for (int i = 0; i < 3; ++i) ypr[i] = o(i); }

}  // extern "C"

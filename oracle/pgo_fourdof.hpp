// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see pgo_core.hpp).
//
// CPU restatement of the reference's alternative edge functors — the family its build keeps switched off
// (call sites commented out at src/PoseGraphSLAM.cpp:1551,1630; `__USE_YPR_REP` is defined nowhere):
//   * src/CeresResidues.h:226-243   R2ypr<T> (degrees)
//   * src/CeresResidues.h:252-335   FourDOFError                          AutoDiffCostFunction<.,6,4,3,4,3>
//   * src/CeresResidues.h:338-425   FourDOFErrorWithSwitchingConstraints  <.,7,4,3,4,3,1>
//   * src/CeresResidues.h:429-456   NormalizeAngle, AngleLocalParameterization (AutoDiffLocalParameterization<.,1,1>)
//   * src/CeresResidues.h:458-497   YawPitchRollToRotationMatrix, RotationMatrixTranspose, RotationMatrixRotatePoint
//   * src/CeresResidues.h:500-546   QinFourDOFWeightError                 <.,4,1,3,1,3>
// restated literally, templated on the scalar, and differentiated the way Ceres does it: Jets over the AMBIENT
// parameters (4+3+4+3[+1] or 1+3+1+3), then every quaternion block right-multiplied by the 4x3 Plus-Jacobian of
// EigenQuaternionParameterization and every yaw block by the 1x1 Jacobian of the autodiffed angle Plus.
#pragma once
#include "pgo_core.hpp"

namespace pgo {

// src/CeresResidues.h:226-243
template <class T> inline void R2ypr_T(const Mat3<T>& R, T ypr[3]) {
  const T n0 = R.m[0][0], n1 = R.m[1][0], n2 = R.m[2][0];   // n = R.col(0)
  const T o0 = R.m[0][1], o1 = R.m[1][1];                   // o = R.col(1)
  const T a0 = R.m[0][2], a1 = R.m[1][2];                   // a = R.col(2)
  T y = atan2(n1, n0);
  T p = atan2(-n2, n0 * cos(y) + n1 * sin(y));
  T r = atan2(a0 * sin(y) - a1 * cos(y), -o0 * sin(y) + o1 * cos(y));
  ypr[0] = y / T(M_PI) * T(180.0); ypr[1] = p / T(M_PI) * T(180.0); ypr[2] = r / T(M_PI) * T(180.0);
}

// src/CeresResidues.h:252-335.  The observation enters as (q_obs, t_obs), q_obs = Quaterniond(c1_T_c2 rotation) (:260).
struct FourDOFError {
  Quat<double> oq; Vec3<double> ot; double weight;
  template <class T>
  bool operator()(const T* q1, const T* t1, const T* q2, const T* t2, T* res) const {
    Vec3<T> p_1{t1[0], t1[1], t1[2]};
    Quat<T> q_1{q1[0], q1[1], q1[2], q1[3]};
    Vec3<T> p_2{t2[0], t2[1], t2[2]};
    Quat<T> q_2{q2[0], q2[1], q2[2], q2[3]};
    Quat<T> q_1_inverse = qconj(q_1);
    Quat<T> q_12_estimated = qmul(q_1_inverse, q_2);
    Vec3<T> p_12_estimated = qrot(q_1_inverse, Vec3<T>{p_2.x - p_1.x, p_2.y - p_1.y, p_2.z - p_1.z});
    Quat<T> obs_q{T(oq.x), T(oq.y), T(oq.z), T(oq.w)};
    Quat<T> delta_q = qmul(qconj(q_12_estimated), obs_q);
    Vec3<T> delta_t = qrot(qconj(q_12_estimated),
                           Vec3<T>{T(ot.x) - p_12_estimated.x, T(ot.y) - p_12_estimated.y, T(ot.z) - p_12_estimated.z});
    res[0] = delta_t.x; res[1] = delta_t.y; res[2] = delta_t.z;
    Mat3<T> delta_Rot = qtoR(delta_q);
    T delta_ypr[3]; R2ypr_T(delta_Rot, delta_ypr);
    res[3] = T(4.) * delta_ypr[0]; res[4] = T(10.) * delta_ypr[1]; res[5] = T(10.) * delta_ypr[2];
    for (int i = 0; i < 6; ++i) res[i] = res[i] * T(weight);
    return true;
  }
};

// src/CeresResidues.h:338-425.  `weight` is stored but not applied (:393).
struct FourDOFErrorWithSwitchingConstraints {
  Quat<double> oq; Vec3<double> ot; double weight;
  template <class T>
  bool operator()(const T* q1, const T* t1, const T* q2, const T* t2, const T* sw, T* res) const {
    Vec3<T> p_1{t1[0], t1[1], t1[2]};
    Quat<T> q_1{q1[0], q1[1], q1[2], q1[3]};
    Vec3<T> p_2{t2[0], t2[1], t2[2]};
    Quat<T> q_2{q2[0], q2[1], q2[2], q2[3]};
    Quat<T> q_1_inverse = qconj(q_1);
    Quat<T> q_12_estimated = qmul(q_1_inverse, q_2);
    Vec3<T> p_12_estimated = qrot(q_1_inverse, Vec3<T>{p_2.x - p_1.x, p_2.y - p_1.y, p_2.z - p_1.z});
    Quat<T> obs_q{T(oq.x), T(oq.y), T(oq.z), T(oq.w)};
    Quat<T> delta_q = qmul(qconj(q_12_estimated), obs_q);
    Vec3<T> delta_t = qrot(qconj(q_12_estimated),
                           Vec3<T>{T(ot.x) - p_12_estimated.x, T(ot.y) - p_12_estimated.y, T(ot.z) - p_12_estimated.z});
    res[0] = delta_t.x; res[1] = delta_t.y; res[2] = delta_t.z;
    res[6] = T(1.0) * (T(1.0) - sw[0]);
    Mat3<T> delta_Rot = qtoR(delta_q);
    T delta_ypr[3]; R2ypr_T(delta_Rot, delta_ypr);
    res[3] = T(4.) * delta_ypr[0]; res[4] = T(10.0) * delta_ypr[1]; res[5] = T(10.0) * delta_ypr[2];
    T s = sw[0];
    for (int i = 0; i < 7; ++i) res[i] = res[i] * s;
    return true;
  }
};

// src/CeresResidues.h:429-436
template <class T> inline T NormalizeAngle(const T& angle_degrees) {
  if (angle_degrees > T(180.0)) return angle_degrees - T(360.0);
  else if (angle_degrees < T(-180.0)) return angle_degrees + T(360.0);
  else return angle_degrees;
}
// src/CeresResidues.h:440-456 through ceres::AutoDiffLocalParameterization<.,1,1>: Plus and d Plus / d delta at delta = 0
inline double angle_plus(double theta, double delta) { return NormalizeAngle(theta + delta); }
inline double angle_plus_jacobian(double theta) { Jet<1> d(0.0, 0); Jet<1> out = NormalizeAngle(Jet<1>(theta) + d); return out.v[0]; }

// src/CeresResidues.h:458-477
template <class T> inline void YawPitchRollToRotationMatrix(const T yaw, const T pitch, const T roll, T R[9]) {
  T y = yaw / T(180.0) * T(M_PI);
  T p = pitch / T(180.0) * T(M_PI);
  T r = roll / T(180.0) * T(M_PI);
  R[0] = cos(y) * cos(p);
  R[1] = -sin(y) * cos(r) + cos(y) * sin(p) * sin(r);
  R[2] = sin(y) * sin(r) + cos(y) * sin(p) * cos(r);
  R[3] = sin(y) * cos(p);
  R[4] = cos(y) * cos(r) + sin(y) * sin(p) * sin(r);
  R[5] = -cos(y) * sin(r) + sin(y) * sin(p) * cos(r);
  R[6] = -sin(p);
  R[7] = cos(p) * sin(r);
  R[8] = cos(p) * cos(r);
}
// src/CeresResidues.h:479-490
template <class T> inline void RotationMatrixTranspose(const T R[9], T inv_R[9]) {
  inv_R[0] = R[0]; inv_R[1] = R[3]; inv_R[2] = R[6];
  inv_R[3] = R[1]; inv_R[4] = R[4]; inv_R[5] = R[7];
  inv_R[6] = R[2]; inv_R[7] = R[5]; inv_R[8] = R[8];
}
// src/CeresResidues.h:492-497
template <class T> inline void RotationMatrixRotatePoint(const T R[9], const T t[3], T r_t[3]) {
  r_t[0] = R[0] * t[0] + R[1] * t[1] + R[2] * t[2];
  r_t[1] = R[3] * t[0] + R[4] * t[1] + R[5] * t[2];
  r_t[2] = R[6] * t[0] + R[7] * t[1] + R[8] * t[2];
}

// src/CeresResidues.h:500-546
struct QinFourDOFWeightError {
  double t_x, t_y, t_z, relative_yaw, pitch_i, roll_i, weight = 1;
  template <class T>
  bool operator()(const T* yaw_i, const T* ti, const T* yaw_j, const T* tj, T* residuals) const {
    T t_w_ij[3];
    t_w_ij[0] = tj[0] - ti[0]; t_w_ij[1] = tj[1] - ti[1]; t_w_ij[2] = tj[2] - ti[2];
    T w_R_i[9];
    YawPitchRollToRotationMatrix(yaw_i[0], T(pitch_i), T(roll_i), w_R_i);
    T i_R_w[9];
    RotationMatrixTranspose(w_R_i, i_R_w);
    T t_i_ij[3];
    RotationMatrixRotatePoint(i_R_w, t_w_ij, t_i_ij);
    residuals[0] = (t_i_ij[0] - T(t_x)) * T(weight);
    residuals[1] = (t_i_ij[1] - T(t_y)) * T(weight);
    residuals[2] = (t_i_ij[2] - T(t_z)) * T(weight);
    residuals[3] = NormalizeAngle((yaw_j[0] - yaw_i[0] - T(relative_yaw))) * T(weight) / T(10.0);
    return true;
  }
};

// ---------------------------------------------------------------------------------- autodiff drivers
// r[6], J[6][12] tangent columns [th1, t1, th2, t2]
inline void eval_fourdof_autodiff(const FourDOFError& f, const double* q1, const double* t1, const double* q2, const double* t2, double* r, double* J) {
  if (!J) { f(q1, t1, q2, t2, r); return; }
  typedef Jet<14> JT;
  JT jq1[4], jt1[3], jq2[4], jt2[3], res[6];
  for (int i = 0; i < 4; ++i) { jq1[i] = JT(q1[i], i); jq2[i] = JT(q2[i], 7 + i); }
  for (int i = 0; i < 3; ++i) { jt1[i] = JT(t1[i], 4 + i); jt2[i] = JT(t2[i], 11 + i); }
  f(jq1, jt1, jq2, jt2, res);
  double P1[4][3], P2[4][3];
  quat_plus_jacobian(q1, P1); quat_plus_jacobian(q2, P2);
  for (int i = 0; i < 6; ++i) {
    r[i] = res[i].a;
    double* Ji = J + 12 * i;
    for (int c = 0; c < 3; ++c) {
      Ji[c] = res[i].v[0] * P1[0][c] + res[i].v[1] * P1[1][c] + res[i].v[2] * P1[2][c] + res[i].v[3] * P1[3][c];
      Ji[3 + c] = res[i].v[4 + c];
      Ji[6 + c] = res[i].v[7] * P2[0][c] + res[i].v[8] * P2[1][c] + res[i].v[9] * P2[2][c] + res[i].v[10] * P2[3][c];
      Ji[9 + c] = res[i].v[11 + c];
    }
  }
}
// r[7], J[7][13] tangent columns [th1, t1, th2, t2, s]
inline void eval_fourdof_switch_autodiff(const FourDOFErrorWithSwitchingConstraints& f, const double* q1, const double* t1, const double* q2,
                                         const double* t2, const double* s, double* r, double* J) {
  if (!J) { f(q1, t1, q2, t2, s, r); return; }
  typedef Jet<15> JT;
  JT jq1[4], jt1[3], jq2[4], jt2[3], js[1], res[7];
  for (int i = 0; i < 4; ++i) { jq1[i] = JT(q1[i], i); jq2[i] = JT(q2[i], 7 + i); }
  for (int i = 0; i < 3; ++i) { jt1[i] = JT(t1[i], 4 + i); jt2[i] = JT(t2[i], 11 + i); }
  js[0] = JT(s[0], 14);
  f(jq1, jt1, jq2, jt2, js, res);
  double P1[4][3], P2[4][3];
  quat_plus_jacobian(q1, P1); quat_plus_jacobian(q2, P2);
  for (int i = 0; i < 7; ++i) {
    r[i] = res[i].a;
    double* Ji = J + 13 * i;
    for (int c = 0; c < 3; ++c) {
      Ji[c] = res[i].v[0] * P1[0][c] + res[i].v[1] * P1[1][c] + res[i].v[2] * P1[2][c] + res[i].v[3] * P1[3][c];
      Ji[3 + c] = res[i].v[4 + c];
      Ji[6 + c] = res[i].v[7] * P2[0][c] + res[i].v[8] * P2[1][c] + res[i].v[9] * P2[2][c] + res[i].v[10] * P2[3][c];
      Ji[9 + c] = res[i].v[11 + c];
    }
    Ji[12] = res[i].v[14];
  }
}
// r[4], J[4][8] tangent columns [yaw_i, t_i, yaw_j, t_j]
inline void eval_qin_autodiff(const QinFourDOFWeightError& f, const double* yaw_i, const double* ti, const double* yaw_j, const double* tj, double* r, double* J) {
  if (!J) { f(yaw_i, ti, yaw_j, tj, r); return; }
  typedef Jet<8> JT;
  JT jyi[1], jti[3], jyj[1], jtj[3], res[4];
  jyi[0] = JT(yaw_i[0], 0); jyj[0] = JT(yaw_j[0], 4);
  for (int i = 0; i < 3; ++i) { jti[i] = JT(ti[i], 1 + i); jtj[i] = JT(tj[i], 5 + i); }
  f(jyi, jti, jyj, jtj, res);
  const double Pi = angle_plus_jacobian(yaw_i[0]), Pj = angle_plus_jacobian(yaw_j[0]);
  for (int i = 0; i < 4; ++i) {
    r[i] = res[i].a;
    double* Ji = J + 8 * i;
    Ji[0] = res[i].v[0] * Pi; Ji[4] = res[i].v[4] * Pj;
    for (int c = 0; c < 3; ++c) { Ji[1 + c] = res[i].v[1 + c]; Ji[5 + c] = res[i].v[5 + c]; }
  }
}

}  // namespace pgo

// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see below).
//
// CPU restatement of the numerical core of mpkuse/solve_keyframe_pose_graph:
//   * src/CeresResidues.h:19-90    SixDOFError
//   * src/CeresResidues.h:96-141   NodePoseRegularization
//   * src/CeresResidues.h:145-222  SixDOFErrorWithSwitchingConstraints
// plus the Eigen 3.3 primitives those functors use (quaternion product, conjugate,
// q*v, toRotationMatrix, Quaternion(Matrix3), 4x4 inverse) and a forward-mode Jet so
// the functors are differentiated the way ceres::AutoDiffCostFunction does it
// (ambient 6x4 / 6x3 blocks, then right-multiplied by the 4x3 Plus-Jacobian of
// ceres::EigenQuaternionParameterization).
//
// "Parity unpinned": the reference ships no test that pins a numerical result of this
// path and Ceres/Eigen are not available in the build container, so this restatement is
// anchored by its own known-answer tests (tests/test_oracle_*.py): functor KATs,
// Jet-autodiff == closed form == central differences, dense cross-checks of the linear
// solve — and, for the TEXT of the functors, by the reference's own src/CeresResidues.h
// compiled unmodified over a stand-in for the Eigen/Ceres API (oracle/shim/,
// oracle/_ref/libref_functors.so, tests/test_reference_functors.py).  Eigen's arithmetic
// and Ceres' minimiser remain restated.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// arm may include or link this directory.  The product (solve_keyframe_pose_graph_b200/)
// never does.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace pgo {

// ----------------------------------------------------------------------------------
// Forward-mode dual number, the equivalent of ceres::Jet<double,N>.
// ----------------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) {
  Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
template <int N> inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }
template <int N> inline Jet<N>& operator-=(Jet<N>& f, const Jet<N>& g) { f = f - g; return f; }
template <int N> inline Jet<N>& operator*=(Jet<N>& f, const Jet<N>& g) { f = f * g; return f; }
template <int N> inline bool operator>(const Jet<N>& f, const Jet<N>& g) { return f.a > g.a; }
template <int N> inline bool operator<(const Jet<N>& f, const Jet<N>& g) { return f.a < g.a; }
template <int N> inline Jet<N> sqrt(const Jet<N>& f) {
  Jet<N> h; h.a = std::sqrt(f.a); const double d = 0.5 / h.a;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * d; return h; }
inline double sqrt(double x) { return std::sqrt(x); }
// trigonometry for the (switched-off) yaw/pitch/roll functors, pgo_fourdof.hpp
template <int N> inline Jet<N> sin(const Jet<N>& f) { Jet<N> h; h.a = std::sin(f.a); const double c = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> inline Jet<N> cos(const Jet<N>& f) { Jet<N> h; h.a = std::cos(f.a); const double s = -std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
// ceres/jet.h: atan2(g, f) = atan2(g.a, f.a), derivative (-g.a f.v + f.a g.v) / (f.a^2 + g.a^2)
template <int N> inline Jet<N> atan2(const Jet<N>& g, const Jet<N>& f) { Jet<N> h; h.a = std::atan2(g.a, f.a);
  const double tmp = 1.0 / (f.a * f.a + g.a * g.a);
  for (int i = 0; i < N; ++i) h.v[i] = tmp * (-g.a * f.v[i] + f.a * g.v[i]); return h; }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline double scalar_of(double x) { return x; }
template <int N> inline double scalar_of(const Jet<N>& x) { return x.a; }

// ----------------------------------------------------------------------------------
// Eigen 3.3 primitives.  Quaternion coefficient order x,y,z,w.
// ----------------------------------------------------------------------------------
template <class T> struct Quat { T x, y, z, w; };
template <class T> struct Vec3 { T x, y, z; };
template <class T> struct Mat3 { T m[3][3]; };
template <class T> struct Mat4 { T m[4][4]; };

template <class T> inline Quat<T> qconj(const Quat<T>& q) { return Quat<T>{-q.x, -q.y, -q.z, q.w}; }

// Eigen::Quaternion operator* (Hamilton product).
template <class T> inline Quat<T> qmul(const Quat<T>& a, const Quat<T>& b) {
  Quat<T> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
template <class T> inline Vec3<T> cross(const Vec3<T>& a, const Vec3<T>& b) {
  return Vec3<T>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen::Quaternion::_transformVector :  uv = 2 (u x v);  v + w uv + u x uv
template <class T> inline Vec3<T> qrot(const Quat<T>& q, const Vec3<T>& v) {
  Vec3<T> u{q.x, q.y, q.z};
  Vec3<T> uv = cross(u, v);
  uv.x = uv.x + uv.x; uv.y = uv.y + uv.y; uv.z = uv.z + uv.z;
  Vec3<T> uuv = cross(u, uv);
  return Vec3<T>{v.x + q.w * uv.x + uuv.x, v.y + q.w * uv.y + uuv.y, v.z + q.w * uv.z + uuv.z};
}
// Eigen::QuaternionBase::toRotationMatrix
template <class T> inline Mat3<T> qtoR(const Quat<T>& q) {
  const T tx = T(2.0) * q.x, ty = T(2.0) * q.y, tz = T(2.0) * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const T txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const T tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3<T> R;
  R.m[0][0] = T(1.0) - (tyy + tzz); R.m[0][1] = txy - twz; R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz; R.m[1][1] = T(1.0) - (txx + tzz); R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy; R.m[2][1] = tyz + twx; R.m[2][2] = T(1.0) - (txx + tyy);
  return R;
}
// Eigen::internal::quaternionbase_assign_impl<Other,3,3> (Shepperd / Shoemake)
template <class T> inline Quat<T> qfromR(const Mat3<T>& M) {
  T q[4];  // x,y,z,w
  T t = M.m[0][0] + M.m[1][1] + M.m[2][2];
  if (t > T(0.0)) {
    t = sqrt(t + T(1.0));
    q[3] = T(0.5) * t;
    t = T(0.5) / t;
    q[0] = (M.m[2][1] - M.m[1][2]) * t;
    q[1] = (M.m[0][2] - M.m[2][0]) * t;
    q[2] = (M.m[1][0] - M.m[0][1]) * t;
  } else {
    int i = 0;
    if (M.m[1][1] > M.m[0][0]) i = 1;
    if (M.m[2][2] > M.m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(M.m[i][i] - M.m[j][j] - M.m[k][k] + T(1.0));
    q[i] = T(0.5) * t;
    t = T(0.5) / t;
    q[3] = (M.m[k][j] - M.m[j][k]) * t;
    q[j] = (M.m[j][i] + M.m[i][j]) * t;
    q[k] = (M.m[k][i] + M.m[i][k]) * t;
  }
  return Quat<T>{q[0], q[1], q[2], q[3]};
}

// Generic 4x4 inverse by cofactors (Eigen's fixed-size 4x4 path is cofactor based too).
template <class T> inline Mat4<T> inv4(const Mat4<T>& A) {
  const T* a = &A.m[0][0];
  T inv[16];
  inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
  inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
  inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
  inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
  inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
  inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
  inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
  inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
  inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
  inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
  inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
  inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
  inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
  inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
  inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
  inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
  T det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
  T idet = T(1.0) / det;
  Mat4<T> R;
  for (int i = 0; i < 16; ++i) (&R.m[0][0])[i] = inv[i] * idet;
  return R;
}
template <class T> inline Mat4<T> mul4(const Mat4<T>& A, const Mat4<T>& B) {
  Mat4<T> C;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      T s = A.m[i][0] * B.m[0][j];
      for (int k = 1; k < 4; ++k) s = s + A.m[i][k] * B.m[k][j];
      C.m[i][j] = s;
    }
  return C;
}

// PoseManipUtils::raw_xyzw_to_eigenmat  (src/utils/PoseManipUtils.cpp:61-72)
inline Mat4<double> pose_to_mat4(const double* q_xyzw, const double* t) {
  Mat3<double> R = qtoR(Quat<double>{q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]});
  Mat4<double> M;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) M.m[i][j] = 0.0;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M.m[i][j] = R.m[i][j];
  M.m[0][3] = t[0]; M.m[1][3] = t[1]; M.m[2][3] = t[2]; M.m[3][3] = 1.0;
  return M;
}
// PoseManipUtils::eigenmat_to_raw_xyzw  (src/utils/PoseManipUtils.cpp:87-98)
inline void mat4_to_pose(const Mat4<double>& M, double* q_xyzw, double* t) {
  Mat3<double> R;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.m[i][j] = M.m[i][j];
  Quat<double> q = qfromR(R);
  q_xyzw[0] = q.x; q_xyzw[1] = q.y; q_xyzw[2] = q.z; q_xyzw[3] = q.w;
  t[0] = M.m[0][3]; t[1] = M.m[1][3]; t[2] = M.m[2][3];
}
// PoseManipUtils::R2ypr (src/utils/PoseManipUtils.cpp:143-158) — DEGREES.
inline void R2ypr_deg(const Mat4<double>& M, double ypr[3]) {
  const double n0 = M.m[0][0], n1 = M.m[1][0], n2 = M.m[2][0];
  const double o0 = M.m[0][1], o1 = M.m[1][1];
  const double a0 = M.m[0][2], a1 = M.m[1][2];
  const double y = std::atan2(n1, n0);
  const double p = std::atan2(-n2, n0 * std::cos(y) + n1 * std::sin(y));
  const double r = std::atan2(a0 * std::sin(y) - a1 * std::cos(y), -o0 * std::sin(y) + o1 * std::cos(y));
  ypr[0] = y / M_PI * 180.0; ypr[1] = p / M_PI * 180.0; ypr[2] = r / M_PI * 180.0;
}

// ----------------------------------------------------------------------------------
// The three live functors, restated literally (templated on the scalar, as Ceres
// instantiates them with double and with Jets).
// ----------------------------------------------------------------------------------

// src/CeresResidues.h:19-90.  Observation given as (q_obs xyzw, t_obs) — the reference's
// constructor (:22-28) converts the 4x4 with Quaterniond(Matrix3d); callers here do that
// conversion with qfromR() before constructing.
struct SixDOFError {
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
    res[3] = T(2.0) * delta_q.x; res[4] = T(2.0) * delta_q.y; res[5] = T(2.0) * delta_q.z;
    T s = T(1.0);  // dynamic covariance scaling is hard-wired off (:63-66)
    for (int i = 0; i < 6; ++i) res[i] = res[i] * (s * T(weight));
    return true;
  }
};

// src/CeresResidues.h:145-222.  `weight` is stored but not applied (:198).
struct SixDOFErrorWithSwitchingConstraints {
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
    res[3] = T(2.0) * delta_q.x; res[4] = T(2.0) * delta_q.y; res[5] = T(2.0) * delta_q.z;
    res[6] = T(1.0) * (T(1.0) - sw[0]);
    T s = sw[0];
    for (int i = 0; i < 7; ++i) res[i] = res[i] * s;
    return true;
  }
};

// src/CeresResidues.h:96-141.
struct NodePoseRegularization {
  Mat4<double> nodepose; double weight;
  template <class T>
  bool operator()(const T* q1, const T* t1, T* res) const {
    Quat<T> q_1{q1[0], q1[1], q1[2], q1[3]};
    Mat3<T> Rq = qtoR(q_1);
    Mat4<T> npose;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) npose.m[i][j] = T(i == j ? 1.0 : 0.0);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) npose.m[i][j] = Rq.m[i][j];
    npose.m[0][3] = t1[0]; npose.m[1][3] = t1[1]; npose.m[2][3] = t1[2];
    Mat4<T> f;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) f.m[i][j] = T(nodepose.m[i][j]);
    Mat4<T> delta = mul4(inv4(f), npose);
    Mat3<T> R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.m[i][j] = delta.m[i][j];
    Quat<T> delta_q = qfromR(R);
    res[0] = T(weight) * delta.m[0][3]; res[1] = T(weight) * delta.m[1][3]; res[2] = T(weight) * delta.m[2][3];
    res[3] = T(weight) * T(2.0) * delta_q.x; res[4] = T(weight) * T(2.0) * delta_q.y; res[5] = T(weight) * T(2.0) * delta_q.z;
    return true;
  }
};

// ----------------------------------------------------------------------------------
// ceres::EigenQuaternionParameterization [CERES-UPSTREAM 1.12-1.14, local_parameterization.cc]
//   Plus(x, d):  x+ = [sin|d| d/|d| ; cos|d|] (x)  x     (left multiply; storage x,y,z,w)
//   ComputeJacobian rows (x,y,z,w):  [ w, z,-y ; -z, w, x ; y,-x, w ; -x,-y,-z ]
// ----------------------------------------------------------------------------------
inline void quat_plus(const double* x, const double* d, double* xp) {
  const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n > 0.0) {
    const double s = std::sin(n) / n;
    Quat<double> dq{s * d[0], s * d[1], s * d[2], std::cos(n)};
    Quat<double> r = qmul(dq, Quat<double>{x[0], x[1], x[2], x[3]});
    xp[0] = r.x; xp[1] = r.y; xp[2] = r.z; xp[3] = r.w;
  } else {
    xp[0] = x[0]; xp[1] = x[1]; xp[2] = x[2]; xp[3] = x[3];
  }
}
inline void quat_plus_jacobian(const double* x, double J[4][3]) {
  J[0][0] = x[3];  J[0][1] = x[2];  J[0][2] = -x[1];
  J[1][0] = -x[2]; J[1][1] = x[3];  J[1][2] = x[0];
  J[2][0] = x[1];  J[2][1] = -x[0]; J[2][2] = x[3];
  J[3][0] = -x[0]; J[3][1] = -x[1]; J[3][2] = -x[2];
}

// ----------------------------------------------------------------------------------
// Autodiff evaluation in Ceres' order: ambient Jacobian blocks from Jets, then
// (6x4)*(4x3) per quaternion block.  Tangent column order per pose: [dtheta(3), dt(3)].
// ----------------------------------------------------------------------------------
// Odometry / SixDOFError: r[6], J[6][12] cols = [th_c1, t_c1, th_c2, t_c2].
inline void eval_sixdof_autodiff(const SixDOFError& f, const double* q1, const double* t1, const double* q2,
                                 const double* t2, double* r, double* J /*6x12 row-major or null*/) {
  if (!J) { f(q1, t1, q2, t2, r); return; }
  typedef Jet<14> JT;
  JT jq1[4], jt1[3], jq2[4], jt2[3], res[6];
  for (int i = 0; i < 4; ++i) jq1[i] = JT(q1[i], i);
  for (int i = 0; i < 3; ++i) jt1[i] = JT(t1[i], 4 + i);
  for (int i = 0; i < 4; ++i) jq2[i] = JT(q2[i], 7 + i);
  for (int i = 0; i < 3; ++i) jt2[i] = JT(t2[i], 11 + i);
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
// Loop / switching: r[7], J[7][13] cols = [th_c1, t_c1, th_c2, t_c2, s].
inline void eval_switch_autodiff(const SixDOFErrorWithSwitchingConstraints& f, const double* q1, const double* t1,
                                 const double* q2, const double* t2, const double* s, double* r, double* J) {
  if (!J) { f(q1, t1, q2, t2, s, r); return; }
  typedef Jet<15> JT;
  JT jq1[4], jt1[3], jq2[4], jt2[3], js[1], res[7];
  for (int i = 0; i < 4; ++i) jq1[i] = JT(q1[i], i);
  for (int i = 0; i < 3; ++i) jt1[i] = JT(t1[i], 4 + i);
  for (int i = 0; i < 4; ++i) jq2[i] = JT(q2[i], 7 + i);
  for (int i = 0; i < 3; ++i) jt2[i] = JT(t2[i], 11 + i);
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
// Regulariser: r[6], J[6][6] cols = [th, t].
inline void eval_reg_autodiff(const NodePoseRegularization& f, const double* q1, const double* t1, double* r, double* J) {
  if (!J) { f(q1, t1, r); return; }
  typedef Jet<7> JT;
  JT jq1[4], jt1[3], res[6];
  for (int i = 0; i < 4; ++i) jq1[i] = JT(q1[i], i);
  for (int i = 0; i < 3; ++i) jt1[i] = JT(t1[i], 4 + i);
  f(jq1, jt1, res);
  double P1[4][3];
  quat_plus_jacobian(q1, P1);
  for (int i = 0; i < 6; ++i) {
    r[i] = res[i].a;
    double* Ji = J + 6 * i;
    for (int c = 0; c < 3; ++c) {
      Ji[c] = res[i].v[0] * P1[0][c] + res[i].v[1] * P1[1][c] + res[i].v[2] * P1[2][c] + res[i].v[3] * P1[3][c];
      Ji[3 + c] = res[i].v[4 + c];
    }
  }
}

// ----------------------------------------------------------------------------------
// Closed-form tangent Jacobians.  Cross-checked against the autodiff path
// by tests/test_oracle_residuals.py; used by the "best-effort CPU" variant B.
// ----------------------------------------------------------------------------------
// e = [R2^T (R1 t_o - t2 + t1) ; 2 vec(q2* (x) q1 (x) q_o)],  Je (6x12)
inline void sixdof_closed_form(const double* q1, const double* t1, const double* q2, const double* t2,
                               const Quat<double>& oq, const Vec3<double>& ot, double e[6], double Je[6][12]) {
  Quat<double> Q1{q1[0], q1[1], q1[2], q1[3]}, Q2{q2[0], q2[1], q2[2], q2[3]};
  Mat3<double> R1 = qtoR(Q1), R2 = qtoR(Q2);
  double a[3];  // R1 t_o
  for (int i = 0; i < 3; ++i) a[i] = R1.m[i][0] * ot.x + R1.m[i][1] * ot.y + R1.m[i][2] * ot.z;
  double v[3] = {a[0] - t2[0] + t1[0], a[1] - t2[1] + t1[1], a[2] - t2[2] + t1[2]};
  for (int i = 0; i < 3; ++i) e[i] = R2.m[0][i] * v[0] + R2.m[1][i] * v[1] + R2.m[2][i] * v[2];
  Quat<double> b = qmul(Q1, oq);
  Quat<double> A = qconj(Q2);
  Quat<double> dq = qmul(A, b);
  e[3] = 2.0 * dq.x; e[4] = 2.0 * dq.y; e[5] = 2.0 * dq.z;
  if (!Je) return;
  auto skew = [](const double* x, double S[3][3]) {
    S[0][0] = 0; S[0][1] = -x[2]; S[0][2] = x[1];
    S[1][0] = x[2]; S[1][1] = 0; S[1][2] = -x[0];
    S[2][0] = -x[1]; S[2][1] = x[0]; S[2][2] = 0;
  };
  double Sa[3][3], Sv[3][3];
  skew(a, Sa); skew(v, Sv);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double m1 = 0, m2 = 0;
      for (int k = 0; k < 3; ++k) { m1 += R2.m[k][i] * Sa[k][j]; m2 += R2.m[k][i] * Sv[k][j]; }
      Je[i][j] = -2.0 * m1;          // d dt / d theta1
      Je[i][3 + j] = R2.m[j][i];     // d dt / d t1  = R2^T
      Je[i][6 + j] = 2.0 * m2;       // d dt / d theta2
      Je[i][9 + j] = -R2.m[j][i];    // d dt / d t2
    }
  // M = (L(q2*) Rm(q1 (x) q_o))[0:3,0:3] = -b_v a_v^T + b_w (a_w I + [a_v]x) - [b_v]x (a_w I + [a_v]x)
  double av[3] = {A.x, A.y, A.z}, bv[3] = {b.x, b.y, b.z};
  double Sav[3][3], Sbv[3][3], G[3][3];
  skew(av, Sav); skew(bv, Sbv);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) G[i][j] = (i == j ? A.w : 0.0) + Sav[i][j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double sg = 0;
      for (int k = 0; k < 3; ++k) sg += Sbv[i][k] * G[k][j];
      const double M = -bv[i] * av[j] + b.w * G[i][j] - sg;
      Je[3 + i][j] = 2.0 * M; Je[3 + i][3 + j] = 0.0; Je[3 + i][6 + j] = -2.0 * M; Je[3 + i][9 + j] = 0.0;
    }
}

inline void eval_sixdof_closed(const SixDOFError& f, const double* q1, const double* t1, const double* q2,
                               const double* t2, double* r, double* J) {
  double e[6], Je[6][12];
  sixdof_closed_form(q1, t1, q2, t2, f.oq, f.ot, e, J ? Je : nullptr);
  for (int i = 0; i < 6; ++i) {
    r[i] = f.weight * e[i];
    if (J) for (int j = 0; j < 12; ++j) J[12 * i + j] = f.weight * Je[i][j];
  }
}
inline void eval_switch_closed(const SixDOFErrorWithSwitchingConstraints& f, const double* q1, const double* t1,
                               const double* q2, const double* t2, const double* sw, double* r, double* J) {
  double e[6], Je[6][12];
  sixdof_closed_form(q1, t1, q2, t2, f.oq, f.ot, e, J ? Je : nullptr);
  const double s = sw[0];
  for (int i = 0; i < 6; ++i) r[i] = s * e[i];
  r[6] = s * (1.0 - s);
  if (J) {
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 12; ++j) J[13 * i + j] = s * Je[i][j]; J[13 * i + 12] = e[i]; }
    for (int j = 0; j < 12; ++j) J[13 * 6 + j] = 0.0;
    J[13 * 6 + 12] = 1.0 - 2.0 * s;
  }
}
// Regulariser closed form: d = q_f* (x) q sign-normalised as Eigen's Quaternion(Matrix3) would.
inline void eval_reg_closed(const NodePoseRegularization& f, const double* q1, const double* t1, double* r, double* J) {
  Mat3<double> Rf; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rf.m[i][j] = f.nodepose.m[i][j];
  Quat<double> qf = qfromR(Rf);
  const double tf[3] = {f.nodepose.m[0][3], f.nodepose.m[1][3], f.nodepose.m[2][3]};
  const double w = f.weight;
  double dt[3] = {t1[0] - tf[0], t1[1] - tf[1], t1[2] - tf[2]};
  for (int i = 0; i < 3; ++i) r[i] = w * (Rf.m[0][i] * dt[0] + Rf.m[1][i] * dt[1] + Rf.m[2][i] * dt[2]);
  Quat<double> A = qconj(qf), Q{q1[0], q1[1], q1[2], q1[3]};
  Quat<double> d = qmul(A, Q);
  // sign that Quaternion(Matrix3) picks: trace = 4w^2-1 > 0 -> w>0 ; else largest |x|,|y|,|z| positive
  double sgn;
  const double n2 = d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
  const double tr = (4.0 * d.w * d.w - n2) / n2;  // trace of R(d/|d|)
  if (tr > 0.0) sgn = d.w >= 0 ? 1.0 : -1.0;
  else {
    const double m0 = d.x * d.x, m1 = d.y * d.y, m2 = d.z * d.z;
    int i = 0; double mi = m0; if (m1 > mi) { i = 1; mi = m1; } if (m2 > mi) { i = 2; }
    const double c = i == 0 ? d.x : (i == 1 ? d.y : d.z);
    sgn = c >= 0 ? 1.0 : -1.0;
  }
  r[3] = w * 2.0 * sgn * d.x; r[4] = w * 2.0 * sgn * d.y; r[5] = w * 2.0 * sgn * d.z;
  if (!J) return;
  // d+ = q_f* (x) dq (x) q  => d vec / d delta = (L(q_f*) Rm(q))[0:3,0:3]
  auto skew = [](const double* x, double S[3][3]) {
    S[0][0] = 0; S[0][1] = -x[2]; S[0][2] = x[1];
    S[1][0] = x[2]; S[1][1] = 0; S[1][2] = -x[0];
    S[2][0] = -x[1]; S[2][1] = x[0]; S[2][2] = 0;
  };
  double av[3] = {A.x, A.y, A.z}, bv[3] = {Q.x, Q.y, Q.z};
  double Sav[3][3], Sbv[3][3], G[3][3];
  skew(av, Sav); skew(bv, Sbv);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) G[i][j] = (i == j ? A.w : 0.0) + Sav[i][j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double sg = 0;
      for (int k = 0; k < 3; ++k) sg += Sbv[i][k] * G[k][j];
      const double M = -bv[i] * av[j] + Q.w * G[i][j] - sg;
      J[6 * i + j] = 0.0; J[6 * i + 3 + j] = w * Rf.m[j][i];
      J[6 * (3 + i) + j] = 2.0 * w * sgn * M; J[6 * (3 + i) + 3 + j] = 0.0;
    }
}

}  // namespace pgo

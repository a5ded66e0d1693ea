// The reference's alternative edge functors (compiled out there) as device math:
//   FourDOFError                          src/CeresResidues.h:252-335   (quaternion blocks, yaw/pitch/roll residual)
//   FourDOFErrorWithSwitchingConstraints  src/CeresResidues.h:338-425
//   QinFourDOFWeightError                 src/CeresResidues.h:500-546   (the __USE_YPR_REP path, PoseGraphSLAM.cpp:1534-1548,1608-1626)
//   AngleLocalParameterization            src/CeresResidues.h:440-456
// Differentiation: forward-mode duals seeded directly in the TANGENT space of the parameter blocks — a quaternion
// block enters with the columns of EigenQuaternionParameterization's 4x3 Plus-Jacobian as its derivative part, which
// is the chain rule Ceres applies after autodiff (ambient 6x4 block times Plus-Jacobian) done in one pass; the angle
// block's Plus is NormalizeAngle(theta + delta) whose Jacobian is 1.  The yaw/pitch/roll extraction goes through
// atan2, so these residuals have no compact closed-form Jacobian worth hand-deriving for a path the reference keeps
// switched off.
//
// PGS_HD lets tests/fourdof_hostcheck.cpp compile this very header for the host (test infrastructure: it checks the
// arithmetic against the oracle without a GPU); libpgs.so only ever instantiates it in __global__ kernels.
#pragma once
#include <math.h>

#ifndef PGS_HD
#define PGS_HD __host__ __device__ __forceinline__
#endif

namespace pgs {
namespace fourdof {

template <int N>
struct Dual {
  double a;
  double v[N];
};

template <int N> PGS_HD Dual<N> dconst(double s) { Dual<N> h; h.a = s;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = 0.0; return h; }
template <int N> PGS_HD Dual<N> dvar(double s, int k) { Dual<N> h = dconst<N>(s); h.v[k] = 1.0; return h; }

template <int N> PGS_HD Dual<N> operator+(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; h.a = f.a + g.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> PGS_HD Dual<N> operator-(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; h.a = f.a - g.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> PGS_HD Dual<N> operator-(const Dual<N>& f) { Dual<N> h; h.a = -f.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> PGS_HD Dual<N> operator*(const Dual<N>& f, const Dual<N>& g) { Dual<N> h; h.a = f.a * g.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> PGS_HD Dual<N> operator*(double s, const Dual<N>& g) { Dual<N> h; h.a = s * g.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = s * g.v[i]; return h; }
template <int N> PGS_HD Dual<N> operator/(const Dual<N>& f, double s) { Dual<N> h; h.a = f.a / s;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] / s; return h; }
template <int N> PGS_HD Dual<N> operator+(const Dual<N>& f, double s) { Dual<N> h = f; h.a += s; return h; }
template <int N> PGS_HD Dual<N> operator-(const Dual<N>& f, double s) { Dual<N> h = f; h.a -= s; return h; }
template <int N> PGS_HD Dual<N> operator-(double s, const Dual<N>& f) { Dual<N> h = -f; h.a += s; return h; }
template <int N> PGS_HD Dual<N> dsin(const Dual<N>& f) { Dual<N> h; double s, c; sincos(f.a, &s, &c); h.a = s;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> PGS_HD Dual<N> dcos(const Dual<N>& f) { Dual<N> h; double s, c; sincos(f.a, &s, &c); h.a = c;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = -s * f.v[i]; return h; }
// atan2(y, x):  d = (x dy - y dx) / (x^2 + y^2)
template <int N> PGS_HD Dual<N> datan2(const Dual<N>& y, const Dual<N>& x) { Dual<N> h; h.a = atan2(y.a, x.a);
  const double inv = 1.0 / (x.a * x.a + y.a * y.a);
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = inv * (x.a * y.v[i] - y.a * x.v[i]); return h; }

template <class T> struct Q { T x, y, z, w; };
template <class T> struct V { T x, y, z; };

template <class T> PGS_HD Q<T> conj(const Q<T>& q) { return Q<T>{-q.x, -q.y, -q.z, q.w}; }
template <class T> PGS_HD Q<T> mul(const Q<T>& a, const Q<T>& b) {   // Hamilton product, Eigen::Quaternion operator*
  Q<T> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
template <class T> PGS_HD V<T> cross(const V<T>& a, const V<T>& b) { return V<T>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
template <class T> PGS_HD V<T> rot(const Q<T>& q, const V<T>& v) {   // Eigen _transformVector: v + w (2 u x v) + u x (2 u x v)
  const V<T> u{q.x, q.y, q.z};
  V<T> uv = cross(u, v);
  uv.x = uv.x + uv.x; uv.y = uv.y + uv.y; uv.z = uv.z + uv.z;
  const V<T> uuv = cross(u, uv);
  return V<T>{v.x + q.w * uv.x + uuv.x, v.y + q.w * uv.y + uuv.y, v.z + q.w * uv.z + uuv.z};
}

// R2ypr (CeresResidues.h:226-243) of delta_q.toRotationMatrix(), degrees.  Only the five entries it reads are formed.
template <int N> PGS_HD void quat_to_ypr_deg(const Q<Dual<N>>& q, Dual<N> ypr[3]) {
  typedef Dual<N> T;
  const T tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const T txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const T tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  const T n0 = 1.0 - (tyy + tzz), n1 = txy + twz, n2 = txz - twy;   // column 0
  const T o0 = txy - twz, o1 = 1.0 - (txx + tzz);                   // column 1, rows 0-1
  const T a0 = txz + twy, a1 = tyz - twx;                           // column 2, rows 0-1
  const T y = datan2(n1, n0);
  const T cy = dcos(y), sy = dsin(y);
  const T p = datan2(-n2, n0 * cy + n1 * sy);
  const T r = datan2(a0 * sy - a1 * cy, o1 * cy - o0 * sy);
  const double k = 180.0 / M_PI;
  ypr[0] = k * y; ypr[1] = k * p; ypr[2] = k * r;
}

// Tangent seed of a quaternion block: value q, derivative columns = Plus-Jacobian rows [w z -y; -z w x; y -x w; -x -y -z]
// placed at dual slots base..base+2 (ceres::EigenQuaternionParameterization::ComputeJacobian).
template <int N> PGS_HD Q<Dual<N>> seed_quat(const double* q, int base) {
  Q<Dual<N>> s{dconst<N>(q[0]), dconst<N>(q[1]), dconst<N>(q[2]), dconst<N>(q[3])};
  s.x.v[base] = q[3];  s.x.v[base + 1] = q[2];  s.x.v[base + 2] = -q[1];
  s.y.v[base] = -q[2]; s.y.v[base + 1] = q[3];  s.y.v[base + 2] = q[0];
  s.z.v[base] = q[1];  s.z.v[base + 1] = -q[0]; s.z.v[base + 2] = q[3];
  s.w.v[base] = -q[0]; s.w.v[base + 1] = -q[1]; s.w.v[base + 2] = -q[2];
  return s;
}
template <int N> PGS_HD V<Dual<N>> seed_vec(const double* t, int base) {
  return V<Dual<N>>{dvar<N>(t[0], base), dvar<N>(t[1], base + 1), dvar<N>(t[2], base + 2)};
}

// FourDOFError (SW = false, N = 12, res[6]) and FourDOFErrorWithSwitchingConstraints (SW = true, N = 13, res[7]).
// Dual slots: [theta1(3), t1(3), theta2(3), t2(3), s].  (CeresResidues.h:268-307 and :355-396.)
template <bool SW, int N>
PGS_HD void four_dof_error(const double* q1, const double* t1, const double* q2, const double* t2, const double* oq, const double* ot,
                           double weight, double sw, Dual<N>* res) {
  typedef Dual<N> T;
  const Q<T> q_1 = seed_quat<N>(q1, 0);  const V<T> p_1 = seed_vec<N>(t1, 3);
  const Q<T> q_2 = seed_quat<N>(q2, 6);  const V<T> p_2 = seed_vec<N>(t2, 9);
  const Q<T> q_1_inverse = conj(q_1);
  const Q<T> q_12_estimated = mul(q_1_inverse, q_2);
  const V<T> p_12_estimated = rot(q_1_inverse, V<T>{p_2.x - p_1.x, p_2.y - p_1.y, p_2.z - p_1.z});
  const Q<T> obs_q{dconst<N>(oq[0]), dconst<N>(oq[1]), dconst<N>(oq[2]), dconst<N>(oq[3])};
  const Q<T> delta_q = mul(conj(q_12_estimated), obs_q);
  const V<T> delta_t = rot(conj(q_12_estimated), V<T>{ot[0] - p_12_estimated.x, ot[1] - p_12_estimated.y, ot[2] - p_12_estimated.z});
  T ypr[3];
  quat_to_ypr_deg<N>(delta_q, ypr);
  res[0] = delta_t.x; res[1] = delta_t.y; res[2] = delta_t.z;
  res[3] = 4.0 * ypr[0]; res[4] = 10.0 * ypr[1]; res[5] = 10.0 * ypr[2];
  if (SW) {
    const T s = dvar<N>(sw, N - 1);
    res[6] = 1.0 - s;                                     // T(1.0) * (T(1.0) - switching_var[0])
#pragma unroll
    for (int i = 0; i < 7; ++i) res[i] = res[i] * s;      // residuals *= s; the weight is not applied (:393)
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i) res[i] = weight * res[i];
  }
}

// NormalizeAngle (CeresResidues.h:429-436): one wrap, strict comparisons, degrees; derivative 1 on every branch.
template <int N> PGS_HD Dual<N> normalize_angle(const Dual<N>& a) {
  if (a.a > 180.0) return a - 360.0;
  if (a.a < -180.0) return a + 360.0;
  return a;
}

// QinFourDOFWeightError (CeresResidues.h:500-546), N = 8, dual slots [yaw_i, t_i(3), yaw_j, t_j(3)], res[4]; angles in degrees.
PGS_HD void qin_four_dof(double yaw_i_, const double* ti_, double yaw_j_, const double* tj_, const double* t_obs, double relative_yaw,
                         double pitch_i, double roll_i, Dual<8>* res) {
  typedef Dual<8> T;
  const T yaw_i = dvar<8>(yaw_i_, 0), yaw_j = dvar<8>(yaw_j_, 4);
  const V<T> ti = seed_vec<8>(ti_, 1), tj = seed_vec<8>(tj_, 5);
  const T tw[3] = {tj.x - ti.x, tj.y - ti.y, tj.z - ti.z};
  // YawPitchRollToRotationMatrix (:458-477)
  const T y = (1.0 / 180.0 * M_PI) * yaw_i;
  const double p = pitch_i / 180.0 * M_PI, r = roll_i / 180.0 * M_PI;
  const T cy = dcos(y), sy = dsin(y);
  const double cp = cos(p), sp = sin(p), cr = cos(r), sr = sin(r);
  T R[9];
  R[0] = cp * cy;              R[1] = (sp * sr) * cy - cr * sy;   R[2] = sr * sy + (sp * cr) * cy;
  R[3] = cp * sy;              R[4] = cr * cy + (sp * sr) * sy;   R[5] = (sp * cr) * sy - sr * cy;
  R[6] = dconst<8>(-sp);       R[7] = dconst<8>(cp * sr);          R[8] = dconst<8>(cp * cr);
  // i_R_w = transpose (:479-490); t_i_ij = i_R_w t_w_ij (:492-497)
  const T l0 = R[0] * tw[0] + R[3] * tw[1] + R[6] * tw[2];
  const T l1 = R[1] * tw[0] + R[4] * tw[1] + R[7] * tw[2];
  const T l2 = R[2] * tw[0] + R[5] * tw[1] + R[8] * tw[2];
  res[0] = l0 - t_obs[0]; res[1] = l1 - t_obs[1]; res[2] = l2 - t_obs[2];      // weight = 1 (:504)
  res[3] = normalize_angle<8>(yaw_j - yaw_i - relative_yaw) / 10.0;
}

}  // namespace fourdof
}  // namespace pgs

// Minimal fixed-size pose algebra for the ROS-free / Eigen-free host side.
//
// The reference passes poses around as Eigen::Matrix4d and converts to/from the optimiser's
// (x,y,z,w | t) storage with PoseManipUtils (reference src/utils/PoseManipUtils.cpp:61-98,
// R2ypr at :143-158).  Eigen is not available in this build, so `pgs::Matrix4d` provides the
// handful of operations the solver front-end needs with Eigen's semantics.  If a maintainer
// builds inside the reference tree, INTEGRATION.md shows the two-line adapter from
// Eigen::Matrix4d.
#pragma once
#include <cmath>
#include <cstring>

namespace pgs {

struct Matrix4d {
  double m[16];  // row-major
  double& operator()(int r, int c) { return m[4 * r + c]; }
  double operator()(int r, int c) const { return m[4 * r + c]; }
  static Matrix4d Identity() {
    Matrix4d I; std::memset(I.m, 0, sizeof(I.m)); I.m[0] = I.m[5] = I.m[10] = I.m[15] = 1.0; return I;
  }
  static Matrix4d Zero() { Matrix4d Z; std::memset(Z.m, 0, sizeof(Z.m)); return Z; }
  Matrix4d operator*(const Matrix4d& B) const {
    Matrix4d C;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double s = 0.0;
        for (int k = 0; k < 4; ++k) s += m[4 * r + k] * B.m[4 * k + c];
        C.m[4 * r + c] = s;
      }
    return C;
  }
  // General 4x4 inverse (Gauss-Jordan with partial pivoting); poses are rigid so this is
  // well conditioned.  The reference calls Eigen's generic Matrix4d::inverse() on poses
  // (PoseGraphSLAM.cpp:1464,1599,1773).
  Matrix4d inverse() const {
    double a[4][8];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { a[r][c] = m[4 * r + c]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
    for (int col = 0; col < 4; ++col) {
      int piv = col;
      for (int r = col + 1; r < 4; ++r) if (std::fabs(a[r][col]) > std::fabs(a[piv][col])) piv = r;
      if (piv != col) for (int c = 0; c < 8; ++c) { double tmp = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = tmp; }
      const double d = 1.0 / a[col][col];
      for (int c = 0; c < 8; ++c) a[col][c] *= d;
      for (int r = 0; r < 4; ++r) if (r != col) { const double f = a[r][col]; if (f != 0.0) for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c]; }
    }
    Matrix4d R;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) R.m[4 * r + c] = a[r][4 + c];
    return R;
  }
};

// Unit quaternion (x,y,z,w) -> rotation part of a 4x4 (Eigen toRotationMatrix form).
inline void quat_to_rot(const double* q, Matrix4d& T) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  T(0, 0) = 1 - 2 * (y * y + z * z); T(0, 1) = 2 * (x * y - z * w);     T(0, 2) = 2 * (x * z + y * w);
  T(1, 0) = 2 * (x * y + z * w);     T(1, 1) = 1 - 2 * (x * x + z * z); T(1, 2) = 2 * (y * z - x * w);
  T(2, 0) = 2 * (x * z - y * w);     T(2, 1) = 2 * (y * z + x * w);     T(2, 2) = 1 - 2 * (x * x + y * y);
}

// raw_xyzw_to_eigenmat (reference PoseManipUtils.cpp:61-72)
inline Matrix4d raw_xyzw_to_mat(const double* quat, const double* t) {
  Matrix4d T = Matrix4d::Zero();
  quat_to_rot(quat, T);
  T(0, 3) = t[0]; T(1, 3) = t[1]; T(2, 3) = t[2]; T(3, 3) = 1.0;
  return T;
}

// eigenmat_to_raw_xyzw (reference PoseManipUtils.cpp:87-98).  Rotation -> quaternion follows
// Eigen's Quaterniond(Matrix3d): positive w when trace > 0, otherwise the largest diagonal
// element picks the positive component.
inline void mat_to_raw_xyzw(const Matrix4d& T, double* quat, double* t) {
  const double tr = T(0, 0) + T(1, 1) + T(2, 2);
  if (tr > 0.0) {
    double s = std::sqrt(tr + 1.0);
    quat[3] = 0.5 * s; s = 0.5 / s;
    quat[0] = (T(2, 1) - T(1, 2)) * s; quat[1] = (T(0, 2) - T(2, 0)) * s; quat[2] = (T(1, 0) - T(0, 1)) * s;
  } else {
    int i = 0;
    if (T(1, 1) > T(0, 0)) i = 1;
    if (T(2, 2) > T(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = std::sqrt(T(i, i) - T(j, j) - T(k, k) + 1.0);
    quat[i] = 0.5 * s; s = 0.5 / s;
    quat[3] = (T(k, j) - T(j, k)) * s; quat[j] = (T(j, i) + T(i, j)) * s; quat[k] = (T(k, i) + T(i, k)) * s;
  }
  t[0] = T(0, 3); t[1] = T(1, 3); t[2] = T(2, 3);
}

// R2ypr (reference PoseManipUtils.cpp:143-158): yaw, pitch, roll in DEGREES.
inline void R2ypr(const Matrix4d& T, double ypr[3]) {
  const double y = std::atan2(T(1, 0), T(0, 0));
  const double p = std::atan2(-T(2, 0), T(0, 0) * std::cos(y) + T(1, 0) * std::sin(y));
  const double r = std::atan2(T(0, 2) * std::sin(y) - T(1, 2) * std::cos(y), -T(0, 1) * std::sin(y) + T(1, 1) * std::cos(y));
  ypr[0] = y / M_PI * 180.0; ypr[1] = p / M_PI * 180.0; ypr[2] = r / M_PI * 180.0;
}
// ypr2R (reference PoseManipUtils.cpp:162-187), degrees in.
inline Matrix4d ypr2R(const double ypr[3]) {
  const double y = ypr[0] / 180.0 * M_PI, p = ypr[1] / 180.0 * M_PI, r = ypr[2] / 180.0 * M_PI;
  const double cy = std::cos(y), sy = std::sin(y), cp = std::cos(p), sp = std::sin(p), cr = std::cos(r), sr = std::sin(r);
  Matrix4d T = Matrix4d::Identity();  // Rz * Ry * Rx
  T(0, 0) = cy * cp; T(0, 1) = cy * sp * sr - sy * cr; T(0, 2) = cy * sp * cr + sy * sr;
  T(1, 0) = sy * cp; T(1, 1) = sy * sp * sr + cy * cr; T(1, 2) = sy * sp * cr - cy * sr;
  T(2, 0) = -sp;     T(2, 1) = cp * sr;                T(2, 2) = cp * cr;
  return T;
}
// rawyprt_to_eigenmat / eigenmat_to_rawyprt (reference PoseManipUtils.cpp:101-141): the (ypr degrees, t) record of the
// __USE_YPR_REP variable store (PoseGraphSLAM.cpp:199-213,228-247) and of QinFourDOFWeightError's constants
// (include/pgs_fourdof.h)
inline Matrix4d rawyprt_to_mat(const double* ypr, const double* t) {
  Matrix4d T = ypr2R(ypr);
  T(0, 3) = t[0]; T(1, 3) = t[1]; T(2, 3) = t[2];
  return T;
}
inline void mat_to_rawyprt(const Matrix4d& T, double* ypr, double* t) {
  R2ypr(T, ypr);
  t[0] = T(0, 3); t[1] = T(1, 3); t[2] = T(2, 3);
}

}  // namespace pgs

// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// A stand-in for the SUBSET of the Eigen 3.3 API that the reference's cost functors use, so that
// /root/reference/src/CeresResidues.h can be compiled UNMODIFIED, from where it lies, in a container that has neither
// Eigen nor Ceres (oracle/Makefile -> oracle/_ref/libref_functors.so; wrapper oracle/ref_functors_capi.cpp).
// What runs is the reference's own functor source — which quaternion is conjugated, what is subtracted from what, the
// residual layout and scaling; what this header supplies is the meaning of the Eigen primitives underneath (Hamilton
// product, q*v, toRotationMatrix, Quaternion(Matrix3), 4x4 inverse, blocks, maps), written from Eigen's documented
// semantics exactly as oracle/pgo_core.hpp states them.  So tests/test_reference_functors.py pins the oracle's
// TRANSCRIPTION of the functors to the reference's text; it does not pin Eigen's or Ceres' arithmetic.
//
// Deliberately small and slow: every matrix is a runtime-sized value with room for 6x6; no expression templates.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace Eigen {

template <class T> struct Dyn;

// assignable view of a rectangular part of a Dyn
template <class T>
struct BlockRef {
  Dyn<T>* m; int i0, j0, r, c;
  inline T& at(int i, int j) const;
  template <class M> BlockRef& assign(const M& src) { for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) at(i, j) = src(i, j); return *this; }
  BlockRef& operator=(const Dyn<T>& src) { return assign(src); }
  BlockRef& operator=(const BlockRef& src) { Dyn<T> tmp = src; return assign(tmp); }
  BlockRef topRows(int n) const { return BlockRef{m, i0, j0, n, c}; }
  T operator()(int i, int j) const { return at(i, j); }
  operator Dyn<T>() const;
};

template <class T>
struct CommaInit {
  Dyn<T>* m; int k;
  CommaInit& operator,(const T& v);
};

template <class T>
struct Dyn {
  int r = 0, c = 0;
  T a[36];
  Dyn() {}
  Dyn(int r_, int c_) : r(r_), c(c_) { for (int i = 0; i < r * c; ++i) a[i] = T(0.0); }
  T& operator()(int i, int j) { return a[i * c + j]; }
  const T& operator()(int i, int j) const { return a[i * c + j]; }
  T& operator()(int i) { return a[i]; }                       // vectors
  const T& operator()(int i) const { return a[i]; }
  int rows() const { return r; }
  int cols() const { return c; }
  BlockRef<T> block(int i0, int j0, int rr, int cc) { return BlockRef<T>{this, i0, j0, rr, cc}; }
  BlockRef<T> topLeftCorner(int rr, int cc) { return BlockRef<T>{this, 0, 0, rr, cc}; }
  BlockRef<T> col(int j) { return BlockRef<T>{this, 0, j, r, 1}; }
  Dyn col(int j) const { Dyn o(r, 1); for (int i = 0; i < r; ++i) o(i, 0) = (*this)(i, j); return o; }
  Dyn topLeftCorner(int rr, int cc) const { Dyn o(rr, cc); for (int i = 0; i < rr; ++i) for (int j = 0; j < cc; ++j) o(i, j) = (*this)(i, j); return o; }
  Dyn topRows(int n) const { return topLeftCorner(n, c); }
  Dyn head(int n) const { Dyn o(n, 1); for (int i = 0; i < n; ++i) o.a[i] = a[i]; return o; }
  Dyn& operator*=(const T& s) { for (int i = 0; i < r * c; ++i) a[i] = a[i] * s; return *this; }
  inline Dyn& operator*=(const Dyn& o);
  CommaInit<T> operator<<(const T& v) { a[0] = v; return CommaInit<T>{this, 1}; }
  template <class U> Dyn<U> cast() const { Dyn<U> o(r, c); for (int i = 0; i < r * c; ++i) o.a[i] = U(a[i]); return o; }
  Dyn transpose() const { Dyn o(c, r); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) o(j, i) = (*this)(i, j); return o; }
  // Eigen's fixed-size 4x4 inverse is cofactor based; same formula as pgo::inv4 (oracle/pgo_core.hpp)
  Dyn inverse() const;
};
template <class T> inline T& BlockRef<T>::at(int i, int j) const { return (*m)(i0 + i, j0 + j); }
template <class T> inline BlockRef<T>::operator Dyn<T>() const { Dyn<T> o(r, c); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) o(i, j) = at(i, j); return o; }
template <class T> inline CommaInit<T>& CommaInit<T>::operator,(const T& v) { m->a[k++] = v; return *this; }

template <class T> Dyn<T> operator+(const Dyn<T>& x, const Dyn<T>& y) { Dyn<T> o(x.r, x.c); for (int i = 0; i < x.r * x.c; ++i) o.a[i] = x.a[i] + y.a[i]; return o; }
template <class T> Dyn<T> operator-(const Dyn<T>& x, const Dyn<T>& y) { Dyn<T> o(x.r, x.c); for (int i = 0; i < x.r * x.c; ++i) o.a[i] = x.a[i] - y.a[i]; return o; }
template <class T> Dyn<T> operator*(const Dyn<T>& x, const Dyn<T>& y) {
  Dyn<T> o(x.r, y.c);
  for (int i = 0; i < x.r; ++i) for (int j = 0; j < y.c; ++j) { T s = x(i, 0) * y(0, j); for (int k = 1; k < x.c; ++k) s = s + x(i, k) * y(k, j); o(i, j) = s; }
  return o;
}
template <class T> inline Dyn<T>& Dyn<T>::operator*=(const Dyn<T>& o) { const Dyn<T> p = (*this) * o; for (int i = 0; i < p.r * p.c; ++i) a[i] = p.a[i]; return *this; }
template <class T> Dyn<T> operator*(const T& s, const Dyn<T>& x) { Dyn<T> o(x.r, x.c); for (int i = 0; i < x.r * x.c; ++i) o.a[i] = s * x.a[i]; return o; }
template <class T> Dyn<T> operator*(const Dyn<T>& x, const T& s) { Dyn<T> o(x.r, x.c); for (int i = 0; i < x.r * x.c; ++i) o.a[i] = x.a[i] * s; return o; }
template <class T> Dyn<T> operator/(const Dyn<T>& x, const T& s) { Dyn<T> o(x.r, x.c); for (int i = 0; i < x.r * x.c; ++i) o.a[i] = x.a[i] / s; return o; }
template <class T> Dyn<T> operator*(const T& s, const BlockRef<T>& b) { return s * Dyn<T>(b); }
// console output only (the reference prints matrices in its debug helpers); not Eigen's IOFormat
template <class T> std::ostream& operator<<(std::ostream& os, const Dyn<T>& m) {
  for (int i = 0; i < m.r; ++i) { for (int j = 0; j < m.c; ++j) os << (j ? " " : "") << m(i, j); if (i + 1 < m.r) os << "\n"; }
  return os;
}

template <class T> Dyn<T> Dyn<T>::inverse() const {
  assert(r == 4 && c == 4);
  const T* m = a; T inv[16];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const T det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  const T idet = T(1.0) / det;
  Dyn o(4, 4);
  for (int i = 0; i < 16; ++i) o.a[i] = inv[i] * idet;
  return o;
}

// Eigen's IOFormat as far as the reference uses it: IOFormat(FullPrecision, DontAlignCols, ", ", "\n") (RawFileIO.h:95-106).
// FullPrecision prints 16 significant digits for double (the sample quoted in src/NodeDataManager.cpp:892-995 shows them).
enum { FullPrecision = -1, StreamPrecision = -2, DontAlignCols = 1 };
struct IOFormat {
  int precision, flags; std::string coeffSeparator, rowSeparator;
  IOFormat(int p = StreamPrecision, int f = 0, const std::string& cs = " ", const std::string& rs = "\n") : precision(p), flags(f), coeffSeparator(cs), rowSeparator(rs) {}
};
template <class T> struct Formatted { const Dyn<T>* m; IOFormat f; };
template <class T> std::ostream& operator<<(std::ostream& os, const Formatted<T>& w) {
  const std::streamsize old = os.precision(w.f.precision == FullPrecision ? 16 : os.precision());
  for (int i = 0; i < w.m->r; ++i) { for (int j = 0; j < w.m->c; ++j) os << (j ? w.f.coeffSeparator : "") << (*w.m)(i, j); if (i + 1 < w.m->r) os << w.f.rowSeparator; }
  os.precision(old);
  return os;
}
template <class Derived> struct MatrixBase {
  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  template <class D = Derived> Formatted<typename D::Scalar> format(const IOFormat& f) const { return Formatted<typename D::Scalar>{&derived(), f}; }
  template <class D = Derived> int rows() const { return static_cast<const D*>(this)->r; }
  template <class D = Derived> int cols() const { return static_cast<const D*>(this)->c; }
};

template <class T, int R, int C>
struct Matrix : Dyn<T>, MatrixBase<Matrix<T, R, C>> {
  typedef T Scalar;
  Matrix() : Dyn<T>(R, C) {}
  explicit Matrix(int) : Dyn<T>(R, C) {}                                   // "Matrix<T,3,1> ypr(3)"
  Matrix(const Dyn<T>& d) : Dyn<T>(R, C) { assert(d.r * d.c == R * C); for (int i = 0; i < R * C; ++i) this->a[i] = d.a[i]; }
  Matrix(const BlockRef<T>& b) : Matrix(Dyn<T>(b)) {}
  static Matrix Identity() { Matrix m; for (int i = 0; i < R && i < C; ++i) m(i, i) = T(1.0); return m; }
  static Matrix Zero() { return Matrix(); }
  static Matrix Zero(int rr, int cc = 1) { Matrix m; assert(rr * cc <= 36); m.r = rr; m.c = cc; return m; }   // the "dynamic" stand-ins
  using Dyn<T>::rows;
  using Dyn<T>::cols;
  Matrix(const T& x, const T& y, const T& z) : Dyn<T>(R, C) { assert(R * C == 3); this->a[0] = x; this->a[1] = y; this->a[2] = z; }
  using Dyn<T>::topLeftCorner;
  template <int P, int Q> BlockRef<T> topLeftCorner() { return BlockRef<T>{this, 0, 0, P, Q}; }
  template <int P, int Q> Matrix<T, P, Q> topLeftCorner() const { return Matrix<T, P, Q>(static_cast<const Dyn<T>&>(*this).topLeftCorner(P, Q)); }
  template <class U> Matrix<U, R, C> cast() const { return Matrix<U, R, C>(Dyn<T>::template cast<U>()); }
  Matrix<T, R, C> inverse() const { return Matrix<T, R, C>(Dyn<T>::inverse()); }
};
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 6, 6> MatrixXd;      // "dynamic" types only appear in declarations the tests never reach
typedef Matrix<float, 6, 6> MatrixXf;
typedef Matrix<double, 6, 1> VectorXd;
typedef Matrix<int, 6, 1> VectorXi;

template <class T>
struct Quaternion {
  T x_, y_, z_, w_;                                                        // Eigen's storage order x,y,z,w
  Quaternion() {}
  Quaternion(const T& w, const T& x, const T& y, const T& z) : x_(x), y_(y), z_(z), w_(w) {}   // constructor order w,x,y,z
  // Eigen::internal::quaternionbase_assign_impl<Other,3,3>
  explicit Quaternion(const Dyn<T>& M) {
    using std::sqrt;
    T q[4];
    T t = M(0, 0) + M(1, 1) + M(2, 2);
    if (t > T(0.0)) {
      t = sqrt(t + T(1.0)); q[3] = T(0.5) * t; t = T(0.5) / t;
      q[0] = (M(2, 1) - M(1, 2)) * t; q[1] = (M(0, 2) - M(2, 0)) * t; q[2] = (M(1, 0) - M(0, 1)) * t;
    } else {
      int i = 0;
      if (M(1, 1) > M(0, 0)) i = 1;
      if (M(2, 2) > M(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = sqrt(M(i, i) - M(j, j) - M(k, k) + T(1.0)); q[i] = T(0.5) * t; t = T(0.5) / t;
      q[3] = (M(k, j) - M(j, k)) * t; q[j] = (M(j, i) + M(i, j)) * t; q[k] = (M(k, i) + M(i, k)) * t;
    }
    x_ = q[0]; y_ = q[1]; z_ = q[2]; w_ = q[3];
  }
  explicit Quaternion(const BlockRef<T>& b) : Quaternion(Dyn<T>(b)) {}
  const T& x() const { return x_; } const T& y() const { return y_; } const T& z() const { return z_; } const T& w() const { return w_; }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Quaternion operator*(const Quaternion& b) const {                        // Hamilton product
    return Quaternion(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                      w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
  // QuaternionBase::_transformVector: uv = 2 (u x v); v + w uv + u x uv
  Matrix<T, 3, 1> operator*(const Dyn<T>& v) const {
    T uv[3] = {y_ * v(2) - z_ * v(1), z_ * v(0) - x_ * v(2), x_ * v(1) - y_ * v(0)};
    for (int i = 0; i < 3; ++i) uv[i] = uv[i] + uv[i];
    const T uuv[3] = {y_ * uv[2] - z_ * uv[1], z_ * uv[0] - x_ * uv[2], x_ * uv[1] - y_ * uv[0]};
    Matrix<T, 3, 1> o;
    for (int i = 0; i < 3; ++i) o(i) = v(i) + w_ * uv[i] + uuv[i];
    return o;
  }
  Matrix<T, 3, 1> vec() const { Matrix<T, 3, 1> o; o(0) = x_; o(1) = y_; o(2) = z_; return o; }
  template <class U> Quaternion<U> cast() const { return Quaternion<U>(U(w_), U(x_), U(y_), U(z_)); }
  Matrix<T, 3, 3> toRotationMatrix() const {                               // QuaternionBase::toRotationMatrix
    const T tx = T(2.0) * x_, ty = T(2.0) * y_, tz = T(2.0) * z_;
    const T twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    Matrix<T, 3, 3> R;
    R(0, 0) = T(1.0) - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = T(1.0) - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = T(1.0) - (txx + tyy);
    return R;
  }
};
typedef Quaternion<double> Quaterniond;

template <class X> struct Map;
// read-only maps: a copy is indistinguishable
template <class T, int R, int C> struct Map<const Matrix<T, R, C>> : Matrix<T, R, C> {
  explicit Map(const T* p) { for (int i = 0; i < R * C; ++i) this->a[i] = p[i]; }
};
template <class T> struct Map<const Quaternion<T>> : Quaternion<T> {
  explicit Map(const T* p) : Quaternion<T>(p[3], p[0], p[1], p[2]) {}       // memory order x,y,z,w
};
// writable map over a column vector: everything writes through
template <class T>
struct PtrBlock {
  T* p; int r;
  PtrBlock& operator=(const Dyn<T>& src) { for (int i = 0; i < r; ++i) p[i] = src.a[i]; return *this; }
};
template <class T, int R> struct Map<Matrix<T, R, 1>> {
  T* p;
  explicit Map(T* p_) : p(p_) {}
  T& operator()(int i) { return p[i]; }
  PtrBlock<T> block(int i0, int j0, int rr, int cc) { assert(j0 == 0 && cc == 1); (void)j0; (void)cc; return PtrBlock<T>{p + i0, rr}; }
  Map& operator*=(const T& s) { for (int i = 0; i < R; ++i) p[i] = p[i] * s; return *this; }
};

}  // namespace Eigen

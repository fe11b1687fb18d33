// =================================================================================================
// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Minimal stand-in for the Eigen3 subset that the
// UNMODIFIED reference sources use (include/common.h:6-9, 3rdPartLib/Sophus/sophus/{so3,se3}.{h,cpp},
// src/map_awareness.cpp, src/map_local.cpp, src/mlmap.cpp, src/rviz_vis.cpp, include/*.h).
//
// Eigen is the reference's one un-vendored arithmetic dependency (CMakeLists.txt:7, system package,
// not present in this image, no network).  Everything here is written from Eigen's documented
// semantics; no expression templates: every operator returns a plain value.  Where the result of a
// floating-point expression depends on Eigen's evaluation ORDER, this file follows Eigen 3.3.x as an
// x86-64 build with the reference's flags (-O3, no -march => SSE2 packets of 2 doubles) evaluates it:
//   * squaredNorm()/norm() of a contiguous double vector: packet-wise partial sums, then the tail
//       3 coefficients: (a0 + a1) + a2          4 coefficients: (a0 + a2) + (a1 + a3)
//   * trace() / dot products of strided coefficients: halving tree   a0 + (a1 + a2)
//   * Quaternion<double> product: the SSE2 kernel of Eigen/src/Geometry/arch/Geometry_SSE.h
//       x = (aw*bx + ay*bz) - (az*by - ax*bw)      y = (aw*by + ay*bw) + (az*bx - ax*bz)
//       z = (aw*bz - ay*bx) + (az*bw + ax*by)      w = (aw*bw - ay*by) - (az*bz + ax*bx)
//   * small matrix products (coefficient based, column-major destination): rows that fall into a
//     full packet are accumulated in order ((l0*r0 + l1*r1) + l2*r2), the odd last row as a tree
//   * _transformVector: uv = qv x v; uv += uv; v + w*uv + qv x uv     normalize(): coeffs / sqrt(squaredNorm)
// No FMA contraction (the reference's build has none either).
// =================================================================================================
#ifndef MLM_REF_MINI_EIGEN_HPP
#define MLM_REF_MINI_EIGEN_HPP

#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <iostream>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3
#define EIGEN_MINOR_VERSION 7

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
const int Dynamic = -1;
typedef std::ptrdiff_t Index;

namespace mini {
template <class T>
struct identity {
  typedef T type;
};
// halving tree of Eigen's redux_novec_unroller
template <class F>
inline auto tree_sum(const F &f, int start, int len) -> decltype(f(0)) {
  if (len == 1) return f(start);
  const int half = len / 2;
  return tree_sum(f, start, half) + tree_sum(f, start + half, len - half);
}
// linear vectorised redux over n contiguous coefficients with packets of P
template <class F>
inline auto packet_sum(const F &f, int n, int P) -> decltype(f(0)) {
  typedef decltype(f(0)) S;
  const int npk = n / P;
  if (npk == 0 || P == 1) return tree_sum(f, 0, n);
  // packets combined by a halving tree (redux_vec_unroller), lanes reduced pairwise low+high (predux)
  std::vector<S> lane(P);
  struct Rec {
    static void run(const F &f, int P, int start, int len, S *out) {
      if (len == 1) {
        for (int l = 0; l < P; l++) out[l] = f(start * P + l);
        return;
      }
      const int half = len / 2;
      std::vector<S> a(P), b(P);
      run(f, P, start, half, a.data());
      run(f, P, start + half, len - half, b.data());
      for (int l = 0; l < P; l++) out[l] = a[l] + b[l];
    }
  };
  Rec::run(f, P, 0, npk, lane.data());
  S res = lane[0];
  if (P == 2) res = lane[0] + lane[1];
  else if (P == 4) res = (lane[0] + lane[2]) + (lane[1] + lane[3]);  // SSE predux<Packet4f>: movehl add, then lane 1
  if (npk * P != n) res = res + tree_sum(f, npk * P, n - npk * P);
  return res;
}
template <class T>
struct packet_size {
  enum { value = 1 };
};
template <>
struct packet_size<double> {
  enum { value = 2 };
};
template <>
struct packet_size<float> {
  enum { value = 4 };
};
template <>
struct packet_size<int> {
  enum { value = 4 };
};
}  // namespace mini

template <class T>
class BlockRef;
template <class T>
class CommaInit;

template <class T, int R, int C, int Opt = 0, int MR = R, int MC = C>
class Matrix {
 public:
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C, SizeAtCompileTime = (R > 0 && C > 0) ? R * C : 1, IsRowMajor = (Opt & RowMajor) ? 1 : 0 };
  typedef T Scalar;
  T d[SizeAtCompileTime];

  static int idx(int i, int j) { return IsRowMajor ? i * C + j : j * R + i; }

  Matrix() {
    for (int i = 0; i < SizeAtCompileTime; i++) d[i] = T();
  }
  Matrix(const T &x, const T &y) {
    static_assert(R * C == 2, "2 coefficients");
    d[0] = x;
    d[1] = y;
  }
  Matrix(const T &x, const T &y, const T &z) {
    static_assert(R * C == 3, "3 coefficients");
    d[0] = x;
    d[1] = y;
    d[2] = z;
  }
  Matrix(const T &x, const T &y, const T &z, const T &w) {
    static_assert(R * C == 4, "4 coefficients");
    d[0] = x;
    d[1] = y;
    d[2] = z;
    d[3] = w;
  }
  explicit Matrix(const T *p) {  // coefficients in storage order
    for (int i = 0; i < SizeAtCompileTime; i++) d[i] = p[i];
  }
  template <int O2>
  Matrix(const Matrix<T, R, C, O2> &o) {
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) (*this)(i, j) = o(i, j);
  }
  template <class U>
  Matrix(const BlockRef<U> &b);
  template <int O2>
  Matrix &operator=(const Matrix<T, R, C, O2> &o) {
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) (*this)(i, j) = o(i, j);
    return *this;
  }
  template <class U>
  Matrix &operator=(const BlockRef<U> &b);

  static int rows() { return R; }
  static int cols() { return C; }
  static int size() { return R * C; }
  T *data() { return d; }
  const T *data() const { return d; }

  T &operator()(int i, int j) { return d[idx(i, j)]; }
  const T &operator()(int i, int j) const { return d[idx(i, j)]; }
  T &coeffRef(int i, int j) { return d[idx(i, j)]; }
  const T &coeff(int i, int j) const { return d[idx(i, j)]; }
  T &operator()(int i) { return d[i]; }
  const T &operator()(int i) const { return d[i]; }
  T &operator[](int i) { return d[i]; }
  const T &operator[](int i) const { return d[i]; }
  T &x() { return d[0]; }
  T &y() { return d[1]; }
  T &z() { return d[2]; }
  T &w() { return d[3]; }
  const T &x() const { return d[0]; }
  const T &y() const { return d[1]; }
  const T &z() const { return d[2]; }
  const T &w() const { return d[3]; }

  Matrix &setZero() {
    for (int i = 0; i < SizeAtCompileTime; i++) d[i] = T(0);
    return *this;
  }
  Matrix &setIdentity() {
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) (*this)(i, j) = i == j ? T(1) : T(0);
    return *this;
  }
  static Matrix Zero() { return Matrix().setZero(); }
  static Matrix Zero(int, int) { return Matrix().setZero(); }
  static Matrix Identity() { return Matrix().setIdentity(); }

  T squaredNorm() const {
    const Matrix &m = *this;
    return mini::packet_sum([&m](int i) { return m.d[i] * m.d[i]; }, SizeAtCompileTime, mini::packet_size<T>::value);
  }
  T norm() const { return std::sqrt(squaredNorm()); }
  void normalize() {
    T z = squaredNorm();
    if (z > T(0)) {
      const T n = std::sqrt(z);
      for (int i = 0; i < SizeAtCompileTime; i++) d[i] = d[i] / n;
    }
  }
  template <int P>
  T lpNorm() const {
    static_assert(P == 1, "only lpNorm<1>");
    T s = T(0);
    for (int i = 0; i < SizeAtCompileTime; i++) s += d[i] < T(0) ? -d[i] : d[i];
    return s;
  }
  T trace() const {
    const Matrix &m = *this;
    return mini::tree_sum([&m](int i) { return m(i, i); }, 0, R);
  }
  Matrix<T, C, R> transpose() const {
    Matrix<T, C, R> t;
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) t(j, i) = (*this)(i, j);
    return t;
  }
  Matrix cross(const Matrix &o) const {
    static_assert(R * C == 3, "cross needs 3 coefficients");
    return Matrix(d[1] * o.d[2] - d[2] * o.d[1], d[2] * o.d[0] - d[0] * o.d[2], d[0] * o.d[1] - d[1] * o.d[0]);
  }

  // ---- block views ----
  BlockRef<T> block(int i, int j, int r, int c) const;
  template <int BR, int BC>
  BlockRef<T> block(int i, int j) const {
    return block(i, j, BR, BC);
  }
  BlockRef<T> row(int i) const { return block(i, 0, 1, C); }
  BlockRef<T> col(int j) const { return block(0, j, R, 1); }
  BlockRef<T> head(int n) const { return C == 1 ? block(0, 0, n, 1) : block(0, 0, 1, n); }
  BlockRef<T> tail(int n) const { return C == 1 ? block(R - n, 0, n, 1) : block(0, C - n, 1, n); }
  template <int N>
  BlockRef<T> head() const {
    return head(N);
  }
  template <int N>
  BlockRef<T> tail() const {
    return tail(N);
  }
  BlockRef<T> topLeftCorner(int r, int c) const { return block(0, 0, r, c); }
  BlockRef<T> topRightCorner(int r, int c) const { return block(0, C - c, r, c); }
  BlockRef<T> bottomLeftCorner(int r, int c) const { return block(R - r, 0, r, c); }
  BlockRef<T> bottomRightCorner(int r, int c) const { return block(R - r, C - c, r, c); }
  template <int BR, int BC>
  BlockRef<T> topLeftCorner() const {
    return block(0, 0, BR, BC);
  }
  template <int BR, int BC>
  BlockRef<T> topRightCorner() const {
    return block(0, C - BC, BR, BC);
  }
  template <int BR, int BC>
  BlockRef<T> bottomLeftCorner() const {
    return block(R - BR, 0, BR, BC);
  }
  template <int BR, int BC>
  BlockRef<T> bottomRightCorner() const {
    return block(R - BR, C - BC, BR, BC);
  }
  CommaInit<T> operator<<(const T &v);

  Matrix &operator+=(const Matrix &o) {
    for (int i = 0; i < SizeAtCompileTime; i++) d[i] = d[i] + o.d[i];
    return *this;
  }
  Matrix &operator-=(const Matrix &o) {
    for (int i = 0; i < SizeAtCompileTime; i++) d[i] = d[i] - o.d[i];
    return *this;
  }
  Matrix &operator*=(const T &s) {
    for (int i = 0; i < SizeAtCompileTime; i++) d[i] = d[i] * s;
    return *this;
  }
  Matrix &operator/=(const T &s) {
    for (int i = 0; i < SizeAtCompileTime; i++) d[i] = d[i] / s;
    return *this;
  }
  bool operator==(const Matrix &o) const {
    for (int i = 0; i < SizeAtCompileTime; i++)
      if (!(d[i] == o.d[i])) return false;
    return true;
  }
  bool operator!=(const Matrix &o) const { return !(*this == o); }
};

// view of a rectangular part of a matrix (also the result of row/col/head/tail/corners); strides in scalars
template <class T>
class BlockRef {
 public:
  T *base;
  int rs, cs, r, c;
  BlockRef(T *b, int rs_, int cs_, int r_, int c_) : base(b), rs(rs_), cs(cs_), r(r_), c(c_) {}
  BlockRef(const BlockRef &) = default;
  T &operator()(int i, int j) const { return base[i * rs + j * cs]; }
  T &operator()(int i) const { return c == 1 ? base[i * rs] : base[i * cs]; }
  T &operator[](int i) const { return (*this)(i); }
  int rows() const { return r; }
  int cols() const { return c; }
  int size() const { return r * c; }
  BlockRef transpose() const { return BlockRef(base, cs, rs, c, r); }
  BlockRef block(int i, int j, int br, int bc) const { return BlockRef(base + i * rs + j * cs, rs, cs, br, bc); }
  BlockRef head(int n) const { return c == 1 ? block(0, 0, n, 1) : block(0, 0, 1, n); }
  BlockRef tail(int n) const { return c == 1 ? block(r - n, 0, n, 1) : block(0, c - n, 1, n); }
  template <int N>
  BlockRef head() const {
    return head(N);
  }
  template <int N>
  BlockRef tail() const {
    return tail(N);
  }
  BlockRef &operator=(const BlockRef &o) {
    for (int i = 0; i < r; i++)
      for (int j = 0; j < c; j++) (*this)(i, j) = o(i, j);
    return *this;
  }
  template <class U>
  BlockRef &operator=(const BlockRef<U> &o) {
    for (int i = 0; i < r; i++)
      for (int j = 0; j < c; j++) (*this)(i, j) = o(i, j);
    return *this;
  }
  template <class U, int R, int C, int O>
  BlockRef &operator=(const Matrix<U, R, C, O> &m) {
    for (int i = 0; i < r; i++)
      for (int j = 0; j < c; j++) (*this)(i, j) = m(i, j);
    return *this;
  }
  T squaredNorm() const {  // contiguous vector segment: same packet order as a plain vector
    const BlockRef &b = *this;
    return mini::packet_sum([&b](int i) { return b(i) * b(i); }, r * c, mini::packet_size<T>::value);
  }
  T norm() const { return std::sqrt(squaredNorm()); }
  CommaInit<T> operator<<(const T &v);
};

// `m << a, b, c, ...;` fills row by row
template <class T>
class CommaInit {
 public:
  BlockRef<T> b;
  int k;
  CommaInit(const BlockRef<T> &b_, const T &first) : b(b_), k(0) { put(first); }
  void put(const T &v) {
    b(k / b.c, k % b.c) = v;
    k++;
  }
  CommaInit &operator,(const T &v) {
    put(v);
    return *this;
  }
};

template <class T, int R, int C, int O, int MR, int MC>
BlockRef<T> Matrix<T, R, C, O, MR, MC>::block(int i, int j, int r, int c) const {
  T *p = const_cast<T *>(d);
  return IsRowMajor ? BlockRef<T>(p + i * C + j, C, 1, r, c) : BlockRef<T>(p + j * R + i, 1, R, r, c);
}
template <class T, int R, int C, int O, int MR, int MC>
CommaInit<T> Matrix<T, R, C, O, MR, MC>::operator<<(const T &v) {
  return CommaInit<T>(block(0, 0, R, C), v);
}
template <class T>
CommaInit<T> BlockRef<T>::operator<<(const T &v) {
  return CommaInit<T>(*this, v);
}
template <class T, int R, int C, int O, int MR, int MC>
template <class U>
Matrix<T, R, C, O, MR, MC>::Matrix(const BlockRef<U> &b) {
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) (*this)(i, j) = b(i, j);
}
template <class T, int R, int C, int O, int MR, int MC>
template <class U>
Matrix<T, R, C, O, MR, MC> &Matrix<T, R, C, O, MR, MC>::operator=(const BlockRef<U> &b) {
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) (*this)(i, j) = b(i, j);
  return *this;
}

// ---- coefficient-wise operators ----
template <class T, int R, int C, int O>
Matrix<T, R, C, O> operator+(const Matrix<T, R, C, O> &a, const Matrix<T, R, C, O> &b) {
  Matrix<T, R, C, O> r;
  for (int i = 0; i < Matrix<T, R, C, O>::SizeAtCompileTime; i++) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <class T, int R, int C, int O>
Matrix<T, R, C, O> operator-(const Matrix<T, R, C, O> &a, const Matrix<T, R, C, O> &b) {
  Matrix<T, R, C, O> r;
  for (int i = 0; i < Matrix<T, R, C, O>::SizeAtCompileTime; i++) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <class T, int R, int C, int O>
Matrix<T, R, C, O> operator-(const Matrix<T, R, C, O> &a) {
  Matrix<T, R, C, O> r;
  for (int i = 0; i < Matrix<T, R, C, O>::SizeAtCompileTime; i++) r.d[i] = -a.d[i];
  return r;
}
template <class T, int R, int C, int O, class U>
Matrix<T, R, C, O> operator+(const Matrix<T, R, C, O> &a, const BlockRef<U> &b) {
  Matrix<T, R, C, O> r;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) r(i, j) = a(i, j) + b(i, j);
  return r;
}
template <class T, int R, int C, int O, class U>
Matrix<T, R, C, O> operator+(const BlockRef<U> &b, const Matrix<T, R, C, O> &a) {
  Matrix<T, R, C, O> r;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) r(i, j) = b(i, j) + a(i, j);
  return r;
}
template <class T, int R, int C, int O, class U>
Matrix<T, R, C, O> operator-(const Matrix<T, R, C, O> &a, const BlockRef<U> &b) {
  Matrix<T, R, C, O> r;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < C; j++) r(i, j) = a(i, j) - b(i, j);
  return r;
}
template <class T, int R, int C, int O>
Matrix<T, R, C, O> operator*(const Matrix<T, R, C, O> &a, const typename mini::identity<T>::type &s) {
  Matrix<T, R, C, O> r;
  for (int i = 0; i < Matrix<T, R, C, O>::SizeAtCompileTime; i++) r.d[i] = a.d[i] * s;
  return r;
}
template <class T, int R, int C, int O>
Matrix<T, R, C, O> operator*(const typename mini::identity<T>::type &s, const Matrix<T, R, C, O> &a) {
  Matrix<T, R, C, O> r;
  for (int i = 0; i < Matrix<T, R, C, O>::SizeAtCompileTime; i++) r.d[i] = s * a.d[i];
  return r;
}
template <class T, int R, int C, int O>
Matrix<T, R, C, O> operator/(const Matrix<T, R, C, O> &a, const typename mini::identity<T>::type &s) {
  Matrix<T, R, C, O> r;
  for (int i = 0; i < Matrix<T, R, C, O>::SizeAtCompileTime; i++) r.d[i] = a.d[i] / s;
  return r;
}
// ---- small products (coefficient based lazy product, column-major destination) ----
template <class T, int R, int K, int C, int O1, int O2>
Matrix<T, R, C> operator*(const Matrix<T, R, K, O1> &a, const Matrix<T, K, C, O2> &b) {
  Matrix<T, R, C> r;
  const int P = mini::packet_size<T>::value;
  const int packed_rows = ((O1 & RowMajor) || R == 1) ? 0 : (R / P) * P;  // rows produced by packet code
  for (int j = 0; j < C; j++)
    for (int i = 0; i < R; i++) {
      if (i < packed_rows) {
        T acc = b(0, j) * a(i, 0);
        for (int k = 1; k < K; k++) acc = b(k, j) * a(i, k) + acc;
        r(i, j) = acc;
      } else {
        r(i, j) = mini::tree_sum([&a, &b, i, j](int k) { return a(i, k) * b(k, j); }, 0, K);
      }
    }
  return r;
}

template <class T, int R, int C, int O>
std::ostream &operator<<(std::ostream &os, const Matrix<T, R, C, O> &m) {
  for (int i = 0; i < R; i++) {
    for (int j = 0; j < C; j++) os << (j ? " " : "") << m(i, j);
    if (i + 1 < R) os << "\n";
  }
  return os;
}
template <class T>
std::ostream &operator<<(std::ostream &os, const BlockRef<T> &m) {
  for (int i = 0; i < m.r; i++) {
    for (int j = 0; j < m.c; j++) os << (j ? " " : "") << m(i, j);
    if (i + 1 < m.r) os << "\n";
  }
  return os;
}

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;

// ---- Quaternion (coefficients stored x, y, z, w like Eigen) ----
template <class T>
class Quaternion {
 public:
  typedef Matrix<T, 3, 1> Vector3;
  typedef Matrix<T, 3, 3> Matrix3;
  Matrix<T, 4, 1> m_coeffs;
  Quaternion() {}
  Quaternion(const T &w, const T &x, const T &y, const T &z) : m_coeffs(x, y, z, w) {}
  explicit Quaternion(const Matrix3 &mat) {  // Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other,3,3>
    T t = mat.trace();
    if (t > T(0)) {
      t = std::sqrt(t + T(1.0));
      w() = T(0.5) * t;
      t = T(0.5) / t;
      x() = (mat.coeff(2, 1) - mat.coeff(1, 2)) * t;
      y() = (mat.coeff(0, 2) - mat.coeff(2, 0)) * t;
      z() = (mat.coeff(1, 0) - mat.coeff(0, 1)) * t;
    } else {
      int i = 0;
      if (mat.coeff(1, 1) > mat.coeff(0, 0)) i = 1;
      if (mat.coeff(2, 2) > mat.coeff(i, i)) i = 2;
      int j = (i + 1) % 3;
      int k = (j + 1) % 3;
      t = std::sqrt(mat.coeff(i, i) - mat.coeff(j, j) - mat.coeff(k, k) + T(1.0));
      m_coeffs[i] = T(0.5) * t;
      t = T(0.5) / t;
      w() = (mat.coeff(k, j) - mat.coeff(j, k)) * t;
      m_coeffs[j] = (mat.coeff(j, i) + mat.coeff(i, j)) * t;
      m_coeffs[k] = (mat.coeff(k, i) + mat.coeff(i, k)) * t;
    }
  }
  T &x() { return m_coeffs[0]; }
  T &y() { return m_coeffs[1]; }
  T &z() { return m_coeffs[2]; }
  T &w() { return m_coeffs[3]; }
  const T &x() const { return m_coeffs[0]; }
  const T &y() const { return m_coeffs[1]; }
  const T &z() const { return m_coeffs[2]; }
  const T &w() const { return m_coeffs[3]; }
  Matrix<T, 4, 1> &coeffs() { return m_coeffs; }
  const Matrix<T, 4, 1> &coeffs() const { return m_coeffs; }
  BlockRef<T> vec() const { return m_coeffs.head(3); }
  Quaternion &setIdentity() {
    m_coeffs = Matrix<T, 4, 1>(T(0), T(0), T(0), T(1));
    return *this;
  }
  static Quaternion Identity() { return Quaternion(T(1), T(0), T(0), T(0)); }
  T squaredNorm() const { return m_coeffs.squaredNorm(); }
  T norm() const { return m_coeffs.norm(); }
  void normalize() { m_coeffs.normalize(); }
  Quaternion normalized() const {
    Quaternion q(*this);
    q.normalize();
    return q;
  }
  Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
  Quaternion inverse() const {
    T n2 = squaredNorm();
    Quaternion c = conjugate();
    c.m_coeffs /= n2;
    return c;
  }
  Quaternion operator*(const Quaternion &b) const {  // quat_product<Architecture::SSE, ..., double>
    const Quaternion &a = *this;
    Quaternion r;
    r.x() = (a.w() * b.x() + a.y() * b.z()) - (a.z() * b.y() - a.x() * b.w());
    r.y() = (a.w() * b.y() + a.y() * b.w()) + (a.z() * b.x() - a.x() * b.z());
    r.z() = (a.w() * b.z() - a.y() * b.x()) + (a.z() * b.w() + a.x() * b.y());
    r.w() = (a.w() * b.w() - a.y() * b.y()) - (a.z() * b.z() + a.x() * b.x());
    return r;
  }
  Quaternion &operator*=(const Quaternion &b) {
    *this = *this * b;
    return *this;
  }
  Vector3 _transformVector(const Vector3 &v) const {
    Vector3 qv = vec();
    Vector3 uv = qv.cross(v);
    uv += uv;
    return v + w() * uv + qv.cross(uv);
  }
  Vector3 operator*(const Vector3 &v) const { return _transformVector(v); }
  Matrix3 toRotationMatrix() const {
    Matrix3 res;
    const T tx = T(2) * x(), ty = T(2) * y(), tz = T(2) * z();
    const T twx = tx * w(), twy = ty * w(), twz = tz * w();
    const T txx = tx * x(), txy = ty * x(), txz = tz * x();
    const T tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    res.coeffRef(0, 0) = T(1) - (tyy + tzz);
    res.coeffRef(0, 1) = txy - twz;
    res.coeffRef(0, 2) = txz + twy;
    res.coeffRef(1, 0) = txy + twz;
    res.coeffRef(1, 1) = T(1) - (txx + tzz);
    res.coeffRef(1, 2) = tyz - twx;
    res.coeffRef(2, 0) = txz - twy;
    res.coeffRef(2, 1) = tyz + twx;
    res.coeffRef(2, 2) = T(1) - (txx + tyy);
    return res;
  }
  Matrix3 matrix() const { return toRotationMatrix(); }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

// scalar * block (e.g. `s * q.vec()`), block * scalar
template <class T>
Matrix<T, 3, 1> operator*(const typename mini::identity<T>::type &s, const BlockRef<T> &b) {
  Matrix<T, 3, 1> r;
  for (int i = 0; i < 3; i++) r[i] = s * b(i);
  return r;
}

}  // namespace Eigen
#endif

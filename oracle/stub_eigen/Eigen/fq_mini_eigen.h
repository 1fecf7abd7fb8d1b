// TEST INFRASTRUCTURE ONLY -- a stand-in for the handful of Eigen features that DecompUtil's headers use, so that the
// REFERENCE's own convex decomposition (thirdparty/DecompROS/DecompUtil/include: ellipsoid_decomp.h, line_segment.h,
// decomp_base.h, ellipsoid.h, polyhedron.h, geometric_utils.h) compiles unmodified from where it lies under /root/reference
// (Eigen is not in this image).  What lives here is plain small-matrix arithmetic -- sums, products, a 3x3 inverse by
// cofactors, quaternion -> rotation matrix -- written the way Eigen evaluates it (pairwise halves for the reductions);
// every line of the decomposition ALGORITHM stays the reference's.  Like oracle/stub_boost for the JPS3D graph search.
// The same stand-in lets the reference's JPS planner layer (jps_planner.cpp, map_util.h: path simplification and ray
// tracing) compile; oracle/stub_ros and oracle/stub_pcl hold the two empty headers map_util.h includes besides.  And the
// reference's solver class itself (faster/src/solverGurobi.cpp with faster_types.hpp: Vector3d states, MatrixXd polytope
// rows, the polynomial solver of getDTInitial) over oracle/stub_gurobi.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <iostream>
#include <memory>
#include <type_traits>
#include <vector>

namespace Eigen
{
const int Dynamic = -1;
const int Infinity = -1;
enum TransformTraits { Isometry = 1, Affine = 2, AffineCompact = 3, Projective = 4 };
template <class T>
using aligned_allocator = std::allocator<T>;
template <typename S>
class Quaternion;

template <typename S, int R, int C>
class Matrix
{
public:
  static constexpr bool kDyn = (R == Dynamic || C == Dynamic);
  typedef typename std::conditional<kDyn, std::vector<S>, std::array<S, (std::size_t)(kDyn ? 1 : R * C)>>::type Store;
  typedef S Scalar;

  Matrix() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C) { init(); }
  // one argument: length of a dynamic vector
  template <class I, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
  explicit Matrix(I n) : r_(R == Dynamic ? (int)n : R), c_(C == Dynamic ? (int)n : C) { init(); }
  // two arguments: (rows, cols) of a dynamic matrix, or the two coefficients of a fixed 2-vector
  template <class A, class B, typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value, int>::type = 0>
  Matrix(const A& a, const B& b) : r_(R == Dynamic ? (int)a : R), c_(C == Dynamic ? (int)b : C)
  {
    init();
    if (!kDyn) { d_[0] = (S)a; d_[1] = (S)b; }
  }
  template <class A, class B, class D, typename std::enable_if<std::is_arithmetic<A>::value, int>::type = 0>
  Matrix(const A& a, const B& b, const D& c) : r_(R), c_(C)
  {
    init();
    d_[0] = (S)a; d_[1] = (S)b; d_[2] = (S)c;
  }
  Matrix(const Quaternion<S>& q);                        // 3x3: the rotation matrix
  // same coefficients in another static shape: Dynamic <-> fixed sizes, or a vector assigned to its transpose (Eigen allows both)
  template <typename S2, int R2, int C2, typename std::enable_if<!(std::is_same<S2, S>::value && R2 == R && C2 == C), int>::type = 0>
  Matrix(const Matrix<S2, R2, C2>& o)
  {
    if (!kDyn) { r_ = R; c_ = C; }
    else if (C == 1) { r_ = o.size(); c_ = 1; }
    else if (R == 1) { r_ = 1; c_ = o.size(); }
    else { r_ = R == Dynamic ? o.rows() : R; c_ = C == Dynamic ? o.cols() : C; }
    init();
    for (int k = 0; k < size() && k < o.size(); k++) d_[(std::size_t)k] = (S)o[k];
  }
  S& x() { return d_[0]; }
  S& y() { return d_[1]; }
  S& z() { return d_[2]; }
  S& w() { return d_[3]; }
  const S& x() const { return d_[0]; }
  const S& y() const { return d_[1]; }
  const S& z() const { return d_[2]; }
  const S& w() const { return d_[3]; }
  // four coefficients (Vector4d)
  template <class A, class B, class D, class E, typename std::enable_if<std::is_arithmetic<A>::value, int>::type = 0>
  Matrix(const A& a, const B& b, const D& c, const E& d) : r_(R), c_(C)
  {
    init();
    d_[0] = (S)a; d_[1] = (S)b; d_[2] = (S)c; d_[3] = (S)d;
  }

  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  S& operator()(int i, int j) { return d_[(std::size_t)(i * c_ + j)]; }
  const S& operator()(int i, int j) const { return d_[(std::size_t)(i * c_ + j)]; }
  S& operator()(int i) { return d_[(std::size_t)i]; }
  const S& operator()(int i) const { return d_[(std::size_t)i]; }
  S& operator[](int i) { return d_[(std::size_t)i]; }
  const S& operator[](int i) const { return d_[(std::size_t)i]; }

  static Matrix Zero() { Matrix m; for (auto& v : m.d_) v = S(0); return m; }
  static Matrix Constant(const S& x) { Matrix m; for (auto& v : m.d_) v = x; return m; }
  static Matrix Identity()
  {
    Matrix m = Zero();
    for (int i = 0; i < m.r_ && i < m.c_; i++) m(i, i) = S(1);
    return m;
  }

  // "m << a, b, c;" fills row by row
  struct Comma
  {
    Matrix* m; int k;
    Comma& operator,(const S& v) { m->d_[(std::size_t)k++] = v; return *this; }
  };
  Comma operator<<(const S& v) { d_[0] = v; return Comma{ this, 1 }; }

  // a row of a matrix as an assignable view (A.row(i) = n: a vector of matching length, either orientation)
  struct RowRef
  {
    Matrix* m; int i;
    template <int R2, int C2>
    RowRef& operator=(const Matrix<S, R2, C2>& v)
    {
      for (int j = 0; j < m->c_; j++) (*m)(i, j) = v(j);
      return *this;
    }
  };
  RowRef row(int i) { return RowRef{ this, i }; }

  Matrix<S, C, R> transpose() const
  {
    Matrix<S, C, R> t = make<C, R>(c_, r_);
    for (int i = 0; i < r_; i++)
      for (int j = 0; j < c_; j++) t(j, i) = (*this)(i, j);
    return t;
  }
  template <typename T>
  Matrix<T, R, C> cast() const
  {
    Matrix<T, R, C> t = Matrix<T, R, C>::sized(r_, c_);
    for (int i = 0; i < r_; i++)
      for (int j = 0; j < c_; j++) t(i, j) = (T)(*this)(i, j);
    return t;
  }
  template <int P>
  S lpNorm() const
  { // only the infinity norm is used (map_util.h:353)
    static_assert(P == Infinity, "only lpNorm<Eigen::Infinity>() is provided");
    S m = S(0);
    for (int i = 0; i < size(); i++) { const S a = d_[(std::size_t)i] < S(0) ? -d_[(std::size_t)i] : d_[(std::size_t)i]; if (a > m) m = a; }
    return m;
  }
  template <int K>
  Matrix<S, K, C> topRows() const
  {
    Matrix<S, K, C> t;
    for (int i = 0; i < K; i++)
      for (int j = 0; j < c_; j++) t(i, j) = (*this)(i, j);
    return t;
  }

  // reductions in pairwise halves, the order of Eigen's unrolled redux for small fixed sizes
  static S halves(const S* p, int n) { return n == 1 ? p[0] : halves(p, n / 2) + halves(p + n / 2, n - n / 2); }
  S squaredNorm() const
  {
    std::vector<S> t((std::size_t)size());
    for (int i = 0; i < size(); i++) t[(std::size_t)i] = d_[(std::size_t)i] * d_[(std::size_t)i];
    return size() ? halves(t.data(), size()) : S(0);
  }
  S norm() const { return std::sqrt(squaredNorm()); }
  Matrix normalized() const
  {
    const S n = norm();
    return n > S(0) ? Matrix(*this) / n : Matrix(*this);
  }
  template <int R2, int C2>
  S dot(const Matrix<S, R2, C2>& o) const
  {
    std::vector<S> t((std::size_t)size());
    for (int i = 0; i < size(); i++) t[(std::size_t)i] = d_[(std::size_t)i] * o(i);
    return size() ? halves(t.data(), size()) : S(0);
  }
  Matrix cross(const Matrix& o) const
  {
    Matrix r;
    r(0) = (*this)(1) * o(2) - (*this)(2) * o(1);
    r(1) = (*this)(2) * o(0) - (*this)(0) * o(2);
    r(2) = (*this)(0) * o(1) - (*this)(1) * o(0);
    return r;
  }
  bool isApprox(const Matrix& o, const S& prec = S(1e-12)) const
  { // |a - b|^2 <= prec^2 min(|a|^2, |b|^2)
    const S a = squaredNorm(), b = o.squaredNorm();
    return (*this - o).squaredNorm() <= prec * prec * (a < b ? a : b);
  }
  S determinant() const
  {
    const Matrix& m = *this;
    if (r_ == 2) return m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
  }
  Matrix inverse() const
  { // cofactors over the determinant (what Eigen does for sizes 2 and 3)
    const Matrix& m = *this;
    Matrix r;
    if (r_ == 2)
    {
      const S id = S(1) / determinant();
      r(0, 0) = m(1, 1) * id; r(0, 1) = -m(0, 1) * id; r(1, 0) = -m(1, 0) * id; r(1, 1) = m(0, 0) * id;
      return r;
    }
    auto cof = [&](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
    };
    const S c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const S det = (c00 * m(0, 0) + c10 * m(1, 0)) + c20 * m(2, 0);
    const S id = S(1) / det;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r(j, i) = cof(i, j) * id;
    return r;
  }

  Matrix operator-() const { Matrix r(*this); for (auto& v : r.d_) v = -v; return r; }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < size(); i++) d_[(std::size_t)i] += o.d_[(std::size_t)i]; return *this; }
  Matrix& operator-=(const Matrix& o) { for (int i = 0; i < size(); i++) d_[(std::size_t)i] -= o.d_[(std::size_t)i]; return *this; }
  template <class T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
  Matrix& operator*=(const T& s) { for (auto& v : d_) v *= (S)s; return *this; }
  template <class T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
  Matrix& operator/=(const T& s) { for (auto& v : d_) v /= (S)s; return *this; }
  friend Matrix operator+(Matrix a, const Matrix& b) { a += b; return a; }
  friend Matrix operator-(Matrix a, const Matrix& b) { a -= b; return a; }
  template <class T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
  friend Matrix operator*(Matrix a, const T& s) { a *= s; return a; }
  template <class T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
  friend Matrix operator*(const T& s, Matrix a) { a *= s; return a; }
  template <class T, typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>
  friend Matrix operator/(Matrix a, const T& s) { a /= s; return a; }
  friend bool operator==(const Matrix& a, const Matrix& b)
  {
    for (int i = 0; i < a.size(); i++) if (!(a.d_[(std::size_t)i] == b.d_[(std::size_t)i])) return false;
    return true;
  }
  friend bool operator!=(const Matrix& a, const Matrix& b) { return !(a == b); }
  friend std::ostream& operator<<(std::ostream& os, const Matrix& m)
  {
    for (int i = 0; i < m.r_; i++)
    {
      for (int j = 0; j < m.c_; j++) os << (j ? " " : "") << m(i, j);
      if (i + 1 < m.r_) os << "\n";
    }
    return os;
  }

  template <int R2, int C2>
  static Matrix<S, R2, C2> make(int r, int c)
  {
    return Matrix<S, R2, C2>::sized(r, c);
  }
  static Matrix sized(int r, int c)
  {
    Matrix m;
    m.r_ = r; m.c_ = c;
    m.init();
    return m;
  }

private:
  void init() { init_store(d_); }
  void init_store(std::vector<S>& v) { v.assign((std::size_t)(r_ * c_), S(0)); }
  template <std::size_t K>
  void init_store(std::array<S, K>&) {}
  int r_, c_;
  Store d_{};
  template <typename, int, int> friend class Matrix;
};

// matrix product; every coefficient is a reduction in pairwise halves
template <typename S, int R, int K, int C>
Matrix<S, R, C> operator*(const Matrix<S, R, K>& a, const Matrix<S, K, C>& b)
{
  Matrix<S, R, C> r = Matrix<S, R, C>::sized(a.rows(), b.cols());
  std::vector<S> t((std::size_t)a.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < b.cols(); j++)
    {
      for (int k = 0; k < a.cols(); k++) t[(std::size_t)k] = a(i, k) * b(k, j);
      r(i, j) = Matrix<S, R, C>::halves(t.data(), a.cols());
    }
  return r;
}

template <typename S>
class Quaternion
{
public:
  Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
  Quaternion(const S& w, const S& x, const S& y, const S& z) : w_(w), x_(x), y_(y), z_(z) {}
  S w() const { return w_; }
  S x() const { return x_; }
  S y() const { return y_; }
  S z() const { return z_; }
  Quaternion operator*(const Quaternion& b) const
  {
    const Quaternion& a = *this;
    return Quaternion(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_, a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                      a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_, a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
  }
  Matrix<S, 3, 3> toRotationMatrix() const
  {
    Matrix<S, 3, 3> r;
    const S tx = S(2) * x_, ty = S(2) * y_, tz = S(2) * z_;
    const S twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_, tyz = tz * y_,
            tzz = tz * z_;
    r(0, 0) = S(1) - (tyy + tzz); r(0, 1) = txy - twz;          r(0, 2) = txz + twy;
    r(1, 0) = txy + twz;          r(1, 1) = S(1) - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy;          r(2, 1) = tyz + twx;          r(2, 2) = S(1) - (txx + tyy);
    return r;
  }
  static Quaternion FromTwoVectors(const Matrix<S, 3, 1>& a, const Matrix<S, 3, 1>& b)
  { // only the visualisation helpers of geometric_utils.h call this (never on the decomposition path)
    const Matrix<S, 3, 1> v0 = a.normalized(), v1 = b.normalized();
    const S c = v0.dot(v1);
    if (c <= S(-1) + S(1e-12))
    { // opposite vectors: any axis perpendicular to a
      Matrix<S, 3, 1> ax = Matrix<S, 3, 1>(1, 0, 0).cross(v0);
      if (ax.norm() < S(1e-6)) ax = Matrix<S, 3, 1>(0, 1, 0).cross(v0);
      ax = ax.normalized();
      return Quaternion(S(0), ax(0), ax(1), ax(2));
    }
    const Matrix<S, 3, 1> ax = v0.cross(v1);
    const S s = std::sqrt((S(1) + c) * S(2)), invs = S(1) / s;
    return Quaternion(s * S(0.5), ax(0) * invs, ax(1) * invs, ax(2) * invs);
  }

private:
  S w_, x_, y_, z_;
};

template <typename S, int R, int C>
Matrix<S, R, C>::Matrix(const Quaternion<S>& q) : r_(R), c_(C)
{
  static_assert(R == 3 && C == 3, "rotation matrix of a quaternion is 3 x 3");
  *this = q.toRotationMatrix();
}
template <typename S>
Matrix<S, 3, 3> operator*(const Matrix<S, 3, 3>& m, const Quaternion<S>& q)
{
  return m * q.toRotationMatrix();
}

template <typename S, int Dim, int Mode>
class Transform
{ // named by a typedef of data_type.h, never used on the decomposition path
};
template <class M>
class SelfAdjointEigenSolver
{ // named by a template of geometric_utils.h that the decomposition never instantiates
public:
  explicit SelfAdjointEigenSolver(const M&) {}
  Matrix<typename M::Scalar, Dynamic, 1> eigenvalues() const { return Matrix<typename M::Scalar, Dynamic, 1>(); }
};
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, 3, 3> Matrix3d;

// unsupported/Eigen/Polynomials: real roots of a polynomial given by its coefficients in ascending order of degree (the
// reference's getDTInitial uses it for a quadratic and a cubic, solverGurobi.cpp:697-735).  Eigen takes the eigenvalues of the
// companion matrix and keeps those whose imaginary part is below 1e-12; this stand-in finds the real roots in closed form
// (discriminant / trigonometric form) and polishes them with Newton steps on the polynomial itself -- the same numbers to
// rounding, by a different route.  What is pinned by compiling the reference is everything AROUND the root finder: which
// polynomials are solved, the float temporaries, MinPositiveElement, the max over the nine times, the division by N.
template <typename S, int Deg>
class PolynomialSolver
{
public:
  template <int R, int C>
  explicit PolynomialSolver(const Matrix<S, R, C>& poly)
  {
    for (int i = 0; i < poly.size(); i++) c_.push_back(poly[i]);
    while (c_.size() > 1 && c_.back() == S(0)) c_.pop_back();       // (Eigen requires a non-zero leading coefficient)
  }
  void realRoots(std::vector<S>& out, const S& = S(1e-12)) const
  {
    out.clear();
    const int deg = (int)c_.size() - 1;
    if (deg == 1) out.push_back(-c_[0] / c_[1]);
    else if (deg == 2)
    {
      const S a = c_[2], b = c_[1], c = c_[0], disc = b * b - S(4) * a * c;
      if (disc >= S(0))
      {
        const S q = -S(0.5) * (b + (b >= S(0) ? std::sqrt(disc) : -std::sqrt(disc)));
        if (q != S(0)) { out.push_back(q / a); out.push_back(c / q); }
        else { out.push_back(S(0)); out.push_back(S(0)); }
      }
    }
    else if (deg == 3)
    {
      const S a = c_[2] / c_[3], b = c_[1] / c_[3], c = c_[0] / c_[3];         // t^3 + a t^2 + b t + c
      const S Q = (a * a - S(3) * b) / S(9), Rr = (S(2) * a * a * a - S(9) * a * b + S(27) * c) / S(54);
      if (Rr * Rr < Q * Q * Q)
      {
        const S th = std::acos(Rr / std::sqrt(Q * Q * Q)), m = -S(2) * std::sqrt(Q);
        const S pi = S(3.14159265358979323846);
        out.push_back(m * std::cos(th / S(3)) - a / S(3));
        out.push_back(m * std::cos((th + S(2) * pi) / S(3)) - a / S(3));
        out.push_back(m * std::cos((th - S(2) * pi) / S(3)) - a / S(3));
      }
      else
      {
        const S A = -(Rr >= S(0) ? S(1) : S(-1)) * std::cbrt(std::fabs(Rr) + std::sqrt(Rr * Rr - Q * Q * Q));
        const S B = A != S(0) ? Q / A : S(0);
        out.push_back((A + B) - a / S(3));
      }
    }
    for (S& r : out)                                                        // Newton polish on the original coefficients
      for (int it = 0; it < 3; it++)
      {
        S f = S(0), df = S(0);
        for (int k = deg; k >= 0; k--) { df = df * r + f; f = f * r + c_[(std::size_t)k]; }
        if (df != S(0) && std::isfinite(f / df)) r -= f / df;
      }
  }

private:
  std::vector<S> c_;
};
}  // namespace Eigen

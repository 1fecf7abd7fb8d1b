// Minimal stand-ins for the Eigen / DecompUtil types that cross the SolverGurobi boundary, used ONLY when the real
// headers are not on the include path (this image has neither Eigen nor DecompUtil).  With the real headers present
// solverGurobi.hpp uses them and this file is not included.
//   Eigen::Vector3d          -> reference faster_types.hpp:79-165 (`state` members)
//   LinearConstraint3D       -> thirdparty/DecompROS/DecompUtil/include/decomp_geometry/polyhedron.h:114-185
#pragma once
#include <cstddef>
#include <vector>

#if __has_include(<Eigen/Dense>)
// Eigen is present but DecompUtil is not: the look-alikes below are expressed with the real Eigen types
#include <Eigen/Dense>
#include <Eigen/StdVector>
#define FQ_COMPAT_HAVE_EIGEN 1
#else
namespace Eigen
{
struct Vector3d
{
  double v[3] = { 0, 0, 0 };
  Vector3d() {}
  Vector3d(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
  static Vector3d Zero() { return Vector3d(); }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double& x() { return v[0]; }
  double& y() { return v[1]; }
  double& z() { return v[2]; }
  double operator()(int i) const { return v[i]; }
  double& operator()(int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
  const Vector3d& transpose() const { return *this; }
};
}  // namespace Eigen
#endif

// vec_Vecf<N> of DecompUtil (decomp_basis/data_type.h:50-80): a vector of N-dimensional points
template <int N>
using Vecf = Eigen::Vector3d;
template <int N>
using vec_Vecf = std::vector<Vecf<N>>;

// F x 3 matrix / F vector with the accessors the solver needs (rows(), (i,j), (i))
struct FqMatX3
{
  std::vector<double> d;   // row-major
  int rows() const { return (int)(d.size() / 3); }
  int cols() const { return 3; }
  double operator()(int i, int j) const { return d[(size_t)3 * i + j]; }
  double& operator()(int i, int j) { return d[(size_t)3 * i + j]; }
  void resize(int r, int) { d.assign((size_t)3 * r, 0.0); }
};
struct FqVecX
{
  std::vector<double> d;
  int rows() const { return (int)d.size(); }
  int size() const { return (int)d.size(); }
  double operator()(int i) const { return d[i]; }
  double& operator()(int i) { return d[i]; }
  double operator[](int i) const { return d[i]; }
  double& operator[](int i) { return d[i]; }
  void resize(int r) { d.assign(r, 0.0); }
};

// A x <= b, outward normals (polyhedron.h:114-185)
struct LinearConstraint3D
{
  LinearConstraint3D() {}
  LinearConstraint3D(const FqMatX3& A, const FqVecX& b) : A_(A), b_(b) {}
  FqMatX3 A() const { return A_; }
  FqVecX b() const { return b_; }
  bool inside(const Eigen::Vector3d& pt) const
  {
    for (int i = 0; i < b_.rows(); i++)
      if (A_(i, 0) * pt.x() + A_(i, 1) * pt.y() + A_(i, 2) * pt.z() - b_(i) > 0) return false;
    return true;
  }
  FqMatX3 A_;
  FqVecX b_;
};

// TEST MOCK of DecompUtil's boundary types, signatures as thirdparty/DecompROS/DecompUtil/include/decomp_geometry/
// polyhedron.h:114-185 (LinearConstraint: A() -> MatDNf<Dim>, b() -> VecDf, inside()) and decomp_basis/data_type.h:50-80
// (Vecf, vec_E, vec_Vecf, MatDNf, VecDf).  Bodies are this mock's own.
#pragma once
#include <Eigen/Dense>
#include <vector>
typedef double decimal_t;
template <int N> using Vecf = Eigen::Matrix<decimal_t, N, 1>;
template <typename T> using vec_E = std::vector<T, Eigen::aligned_allocator<T>>;
template <int N> using vec_Vecf = vec_E<Vecf<N>>;
template <int N> using MatDNf = Eigen::Matrix<decimal_t, Eigen::Dynamic, N>;
typedef Eigen::Matrix<decimal_t, Eigen::Dynamic, 1> VecDf;
template <int Dim>
struct LinearConstraint
{
  LinearConstraint() {}
  LinearConstraint(const MatDNf<Dim>& A, const VecDf& b) : A_(A), b_(b) {}
  bool inside(const Vecf<Dim>& pt)
  {
    for (int i = 0; i < b_.rows(); i++)
    {
      decimal_t d = -b_(i);
      for (int k = 0; k < Dim; k++) d += A_(i, k) * pt(k);
      if (d > 0) return false;
    }
    return true;
  }
  MatDNf<Dim> A() const { return A_; }
  VecDf b() const { return b_; }
  MatDNf<Dim> A_;
  VecDf b_;
};
typedef LinearConstraint<3> LinearConstraint3D;

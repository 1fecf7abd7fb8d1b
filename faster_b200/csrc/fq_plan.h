// Plan tables shared by every candidate with the same (N, force_final).
//
// Normalised time: per axis the spline is the triple integrator driven by the piecewise-constant jerk
// (SURVEY.md section 8a "condensed form", derived from solverGurobi.cpp:359-380,:499-524).  With
//   ubar_t = u_t dt^3,  P_t = p_t,  V_t = v_t dt,  A_t = a_t dt^2      (knot t = start of segment t)
// the recurrence  P' = P + V + A/2 + ubar/6,  V' = V + A + ubar/2,  A' = A + ubar  has no dt in it, so every
// linear functional the model needs is a constant row over ubar.  The final-state equalities C ubar = rhs
// (solverGurobi.cpp:343-356) are eliminated once per (N, force_final):  ubar = Eplus rhs + Z w  with Z an
// orthonormal null-space basis, and the QP becomes  min |w|^2  s.t. inequality rows only.
//
// Y rows (per axis), NY = 6N+1:
//   [0, N]          P_t   position at knot t (t = N: end of the last segment)
//   N+1   + t       V_t   t < N
//   2N+1  + t       A_t
//   3N+1  + t       U_t   ubar_t
//   4N+1  + t       C1_t  Bezier control point 1 of segment t = P_t + V_t/3          (solverGurobi.cpp:840-847)
//   5N+1  + t       C2_t  control point 2 = P_t + 2 V_t/3 + A_t/6                    (solverGurobi.cpp:849-856)
// (control point 0 is P_t, control point 3 is P_{t+1}; :833-838,:858-862)
//
//   Y[y] = T0[y][0..2] . (P0,V0,A0) + T0[y][3..3+ne) . rhs + TZ[y][0..nz) . w
#pragma once
#include <vector>

struct FqPlanHost
{
  int N = 0, force_final = 0;
  int ne = 0;   // eliminated equalities per axis: 3 (whole) or 2 (safe)
  int nz = 0;   // free variables per axis  N - ne
  int NY = 0;   // 6N+1
  std::vector<double> TZ;   // NY x nz
  std::vector<double> T0;   // NY x (3+ne)
  std::vector<double> FT;   // ne x 3 : free response of the terminal rows on (P0,V0,A0)
};

// returns false for unsupported N
bool fq_build_plan(int N, int force_final, FqPlanHost* plan);

// validation scans (fq_host.cpp)
#include <cstddef>
#include <cstdint>
bool fq_scan_all_finite(const double* p, size_t n);
bool fq_scan_all_positive_finite(const double* p, size_t n);
int fq_scan_max_u8(const uint8_t* p, size_t n);

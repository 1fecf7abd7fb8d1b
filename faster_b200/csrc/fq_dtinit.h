// getDTInitial (reference solverGurobi.cpp:659-759) and fillX's per-sample evaluation (:122-168) as ONE source compiled
// for the host (fq_dt_initial, fq_fill_x in fq_host.cpp) and for the device (the chained whole -> safe replan, where the
// safe trajectory's start state R and therefore its time allocations only exist on the GPU, fq_pair.cuh).
//
// Every double operation is spelled through FQ_MUL/FQ_ADD/... so that the device code uses the round-to-nearest
// intrinsics (never contracted into FMAs) and evaluates the SAME expression tree as the host code.  What can still
// differ between the two are acos/cos/cbrt (a few ulp); their results go through three Newton steps and are then
// rounded to float exactly as the reference's float temporaries do (:662-670), which absorbs such differences except
// for results within ~1e-16 relative of a float rounding boundary.  tests/test_pair_gpu.py compares the two.
#pragma once
#include <cmath>

#if defined(__CUDA_ARCH__)
#define FQ_HD __host__ __device__ __forceinline__
#define FQ_MUL(a, b) __dmul_rn((a), (b))
#define FQ_ADD(a, b) __dadd_rn((a), (b))
#define FQ_SUB(a, b) __dsub_rn((a), (b))
#define FQ_DIV(a, b) __ddiv_rn((a), (b))
#define FQ_SQRT(a) __dsqrt_rn((a))
#else
#if defined(__CUDACC__)
#define FQ_HD __host__ __device__ inline
#else
#define FQ_HD inline
#endif
#define FQ_MUL(a, b) ((a) * (b))
#define FQ_ADD(a, b) ((a) + (b))
#define FQ_SUB(a, b) ((a) - (b))
#define FQ_DIV(a, b) ((a) / (b))
#define FQ_SQRT(a) sqrt((a))
#endif

namespace fqdt
{
// real roots of c2 t^2 + c1 t + c0
FQ_HD int roots2(double c0, double c1, double c2, double* r)
{
  const double disc = FQ_SUB(FQ_MUL(c1, c1), FQ_MUL(FQ_MUL(4.0, c2), c0));
  if (disc < 0) return 0;
  const double sq = FQ_SQRT(disc);
  const double q = FQ_MUL(-0.5, FQ_ADD(c1, (c1 >= 0 ? sq : -sq)));
  r[0] = FQ_DIV(q, c2);
  r[1] = q != 0 ? FQ_DIV(c0, q) : 0.0;
  return 2;
}

// real roots of c3 t^3 + c2 t^2 + c1 t + c0 (closed form + three Newton steps)
FQ_HD int roots3(double c0, double c1, double c2, double c3, double* r)
{
  const double a = FQ_DIV(c2, c3), b = FQ_DIV(c1, c3), c = FQ_DIV(c0, c3);
  const double Q = FQ_DIV(FQ_SUB(FQ_MUL(a, a), FQ_MUL(3.0, b)), 9.0);
  const double R = FQ_DIV(FQ_ADD(FQ_SUB(FQ_MUL(FQ_MUL(FQ_MUL(2.0, a), a), a), FQ_MUL(FQ_MUL(9.0, a), b)), FQ_MUL(27.0, c)), 54.0);
  const double RR = FQ_MUL(R, R), QQQ = FQ_MUL(FQ_MUL(Q, Q), Q);
  const double a3 = FQ_DIV(a, 3.0);
  int k = 0;
  if (RR < QQQ)
  {
    const double th = acos(FQ_DIV(R, FQ_SQRT(QQQ))), m = FQ_MUL(-2.0, FQ_SQRT(Q));
    const double two_pi = 6.283185307179586476925286766559;
    r[k++] = FQ_SUB(FQ_MUL(m, cos(FQ_DIV(th, 3.0))), a3);
    r[k++] = FQ_SUB(FQ_MUL(m, cos(FQ_DIV(FQ_ADD(th, two_pi), 3.0))), a3);
    r[k++] = FQ_SUB(FQ_MUL(m, cos(FQ_DIV(FQ_SUB(th, two_pi), 3.0))), a3);
  }
  else
  {
    const double A = -copysign(cbrt(FQ_ADD(fabs(R), FQ_SQRT(FQ_SUB(RR, QQQ)))), R);
    const double B = A != 0 ? FQ_DIV(Q, A) : 0.0;
    r[k++] = FQ_SUB(FQ_ADD(A, B), a3);
    if (RR == QQQ && Q != 0) r[k++] = FQ_SUB(FQ_MUL(-0.5, FQ_ADD(A, B)), a3);
  }
  for (int i = 0; i < k; i++)
    for (int it = 0; it < 3; it++)
    {
      const double t = r[i];
      const double f = FQ_ADD(FQ_MUL(FQ_ADD(FQ_MUL(FQ_ADD(FQ_MUL(c3, t), c2), t), c1), t), c0);
      const double fp = FQ_ADD(FQ_MUL(FQ_ADD(FQ_MUL(FQ_MUL(3.0, c3), t), FQ_MUL(2.0, c2)), t), c1);
      if (fp != 0)
      {
        const double step = FQ_DIV(f, fp);
        if (step - step == 0) r[i] = FQ_SUB(t, step);     // finite step only
      }
    }
  return k;
}

// MinPositiveElement (solverGurobi_utils.hpp:19-32): 0 when there is no positive element
FQ_HD double min_positive(const double* v, int n)
{
  double best = 0;
  bool found = false;
  for (int i = 0; i < n; i++)
    if (v[i] > 0 && (!found || v[i] < best)) { best = v[i]; found = true; }
  return best;
}

// getDTInitial, one axis: the minimum time under the velocity, acceleration and jerk limit alone; float temporaries exactly
// where the reference has them (:662-670).  Returns max(t_v, t_a, t_j) of axis i as float.
FQ_HD float dt_axis(const double* x0, const double* xf, const double* lim, int i)
{
  const double v_max = lim[0], a_max = lim[1], j_max = lim[2];
  const double dp = FQ_SUB(xf[i], x0[i]);
  const float t_v = (float)FQ_DIV(fabs(dp), v_max);                    // :672-674
  const float jerk = (float)FQ_MUL(copysign(1.0, dp), j_max);          // :679-681
  const float accel = (float)FQ_MUL(copysign(1.0, dp), a_max);         // :718-720
  const float a0 = (float)x0[6 + i], v0 = (float)x0[3 + i];            // :682-687
  double r[3];
  int k = roots3(FQ_SUB(x0[i], xf[i]), (double)v0, FQ_DIV((double)a0, 2.0), FQ_DIV((double)jerk, 6.0), r);   // :691-713
  const float t_j = (float)min_positive(r, k);
  k = roots2(FQ_SUB(x0[i], xf[i]), (double)v0, FQ_MUL(0.5, (double)accel), r);                               // :724-746
  const float t_a = (float)min_positive(r, k);
  const float m1 = t_a > t_j ? t_a : t_j;
  return t_v > m1 ? t_v : m1;
}

// the result is the max over axes and limits divided by N, in float (:751)
FQ_HD double dt_from_worst(float worst, int N)
{
  double dt = (double)(worst / (float)N);                                // float / int (:751)
  if (dt > 10000) dt = 0;                                                // :752-756
  return dt;
}

FQ_HD double dt_initial(const double* x0, const double* xf, const double* lim, int N)
{
  float worst = 0;
  for (int i = 0; i < 3; i++)
  {
    const float m2 = dt_axis(x0, xf, lim, i);
    worst = worst > m2 ? worst : m2;
  }
  return dt_from_worst(worst, N);
}

// resetX (:382-388): (int)(N_)*dt_/DC truncated to int, at least 2
FQ_HD int num_samples(int N, double dt, double DC)
{
  const int size = (int)FQ_DIV(FQ_MUL((double)N, dt), DC);
  return size < 2 ? 2 : size;
}

// one sample of fillX (:122-168): polynomial of segment `interval` at local time tau -> pos vel accel jerk (12 doubles)
FQ_HD void eval_sample(const double* x, double tau, double* o)
{
  for (int ax = 0; ax < 3; ax++)
  {
    // x[ax] tau^3 + x[3+ax] tau^2 + x[6+ax] tau + x[9+ax], left to right as the reference writes it (:137-151)
    o[ax] = FQ_ADD(FQ_ADD(FQ_ADD(FQ_MUL(FQ_MUL(FQ_MUL(x[ax], tau), tau), tau), FQ_MUL(FQ_MUL(x[3 + ax], tau), tau)), FQ_MUL(x[6 + ax], tau)), x[9 + ax]);
    o[3 + ax] = FQ_ADD(FQ_ADD(FQ_MUL(FQ_MUL(FQ_MUL(3.0, x[ax]), tau), tau), FQ_MUL(FQ_MUL(2.0, x[3 + ax]), tau)), x[6 + ax]);
    o[6 + ax] = FQ_ADD(FQ_MUL(FQ_MUL(6.0, x[ax]), tau), FQ_MUL(2.0, x[3 + ax]));
    o[9 + ax] = FQ_MUL(6.0, x[ax]);
  }
}
}  // namespace fqdt

// Device side of the chained replan (Faster::replan(), reference faster/src/faster.cpp:406-430,:475,:521-537): for many
// corridors at once
//     whole sweep  ->  genNewTraj selection  ->  R = sample k_safe of fillX  ->  safe sweep from R  ->  selection
// without a host round trip.  The solves are the ordinary batch kernel (fq_kernels_t.cuh); this file holds the small
// kernels between them.  All of them are latency-trivial (one thread / one CTA per corridor, or a grid-stride copy).
#pragma once
#include "fq_dtinit.h"

namespace fqp
{
// dt_base[j] = max(getDTInitial(x0_j, xf_j), 2 DC)  (findDT, solverGurobi.cpp:494-497).  x0 rows with a NaN (a safe
// problem whose whole sweep found nothing) give NaN, which makes every candidate of the problem "not solved".
__global__ void fq_dtbase_kernel(int n_prob, int N, double DC, const double* __restrict__ x0, const double* __restrict__ xf,
                                 const double* __restrict__ lim, double* __restrict__ dt_base)
{ // four lanes per problem: one per axis (the cubic / quadratic root finding is the long part), the fourth idles
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = tid >> 2, ax = tid & 3;
  const bool live = j < n_prob;
  double a[9], b[9], l[3];
  bool ok = live;
  float worst = 0;
  if (live)
  {
    for (int i = 0; i < 9; i++) { a[i] = x0[j * 9 + i]; b[i] = xf[j * 9 + i]; ok = ok && isfinite(a[i]) && isfinite(b[i]); }
    for (int i = 0; i < 3; i++) { l[i] = lim[j * 3 + i]; ok = ok && l[i] > 0 && isfinite(l[i]); }
    if (ok && ax < 3) worst = fqdt::dt_axis(a, b, l, ax);
  }
  worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, 1));
  worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, 2));
  if (!live || ax != 0) return;
  if (!ok) { dt_base[j] = __longlong_as_double(0x7ff8000000000000ll); return; }
  const double dti = fqdt::dt_from_worst(worst, N);
  const double floor2 = FQ_MUL(2.0, DC);
  dt_base[j] = dti > floor2 ? dti : floor2;            // std::max(getDTInitial(), 2 * DC)
}

// candidates of problem j = factors x assignments: dt = factor * dt_base[j] (findDT), sigma from the shared list
__global__ void fq_expand_grid_kernel(int n_prob, int N, int n_fac, int n_sig, const double* __restrict__ factors,
                                      const uint8_t* __restrict__ sig_list, const double* __restrict__ dt_base,
                                      double* __restrict__ dt, uint8_t* __restrict__ sigma, int* __restrict__ cand_ofs)
{
  const long long per = (long long)n_fac * n_sig, total = per * n_prob;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x)
  {
    const int j = (int)(c / per);
    const int r = (int)(c - (long long)j * per);
    const int f = r / n_sig, s = r - f * n_sig;
    dt[c] = FQ_MUL(factors[f], dt_base[j]);
    for (int t = 0; t < N; t++) sigma[c * N + t] = sig_list ? sig_list[(size_t)s * N + t] : (uint8_t)0;
  }
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j <= n_prob; j += (long long)gridDim.x * blockDim.x)
    cand_ofs[j] = (int)(j * per);
}

// genNewTraj's selection (solverGurobi.cpp:445-472) for every problem of a batch: the winner is the feasible candidate
// with the smallest dt (= first feasible factor of the ascending sweep), among those the smallest cost, among those the
// lowest index -- three exact reductions on the ordered bit patterns of the (positive) doubles.  One CTA per problem.
// Also emits the one-candidate-per-problem list that the coefficient re-solve of the winners runs on.

__global__ void __launch_bounds__(128) fq_select_multi_kernel(const FqSelectMultiArgs a)
{
  __shared__ unsigned long long s_dt, s_cost;
  __shared__ int s_idx;
  const int j = blockIdx.x;
  if (j >= a.n_prob) return;
  const int c0 = a.cand_ofs[j], n = a.cand_ofs[j + 1] - c0;
  if (threadIdx.x == 0) { s_dt = ~0ull; s_cost = ~0ull; s_idx = 0x7fffffff; }
  __syncthreads();
  unsigned long long m = ~0ull;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (a.feasible[c0 + i])
    {
      const unsigned long long b = (unsigned long long)__double_as_longlong(a.dt[c0 + i]);
      m = b < m ? b : m;
    }
  if (m != ~0ull) atomicMin(&s_dt, m);
  __syncthreads();
  const unsigned long long dtb = s_dt;
  if (dtb != ~0ull)
  {
    m = ~0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (a.feasible[c0 + i] && (unsigned long long)__double_as_longlong(a.dt[c0 + i]) == dtb)
      {
        const unsigned long long b = (unsigned long long)__double_as_longlong(a.cost[c0 + i]);
        m = b < m ? b : m;
      }
    if (m != ~0ull) atomicMin(&s_cost, m);
    __syncthreads();
    const unsigned long long cb = s_cost;
    int mi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      if (a.feasible[c0 + i] && (unsigned long long)__double_as_longlong(a.dt[c0 + i]) == dtb &&
          (unsigned long long)__double_as_longlong(a.cost[c0 + i]) == cb)
      { mi = i; break; }
    if (mi != 0x7fffffff) atomicMin(&s_idx, mi);
    __syncthreads();
  }
  const int w = dtb != ~0ull ? s_idx : -1;
  if (threadIdx.x == 0)
  {
    a.win_idx[j] = w;
    a.win_cost[j] = w >= 0 ? a.cost[c0 + w] : INFINITY;
    a.win_dt[j] = w >= 0 ? a.dt[c0 + w] : __longlong_as_double(0x7ff8000000000000ll);
    a.win_ofs[j] = j;
    if (j == a.n_prob - 1) a.win_ofs[a.n_prob] = a.n_prob;
  }
  if (a.win_sigma)
    for (int t = threadIdx.x; t < a.N; t += blockDim.x) a.win_sigma[(size_t)j * a.N + t] = w >= 0 && a.sigma ? a.sigma[(size_t)(c0 + w) * a.N + t] : (uint8_t)0;
}

// Between the two sweeps: R = X_temp_[k_safe] of the whole winner (faster.cpp:475; fillX, solverGurobi.cpp:122-168, with the
// reference's accumulated sample time and lagging interval index), k_safe = min(n - 1, (int)(r_fraction * n)) standing in
// for findIndexR (faster.cpp:173-216, out of scope: needs the map).  One thread per problem.

__global__ void fq_pair_mid_kernel(const FqPairMidArgs a)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n_prob) return;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  if (a.win_idx[j] < 0)
  {
    for (int i = 0; i < 9; i++) a.x0_safe[j * 9 + i] = nan;
    a.n_samples[j] = 0; a.k_safe[j] = -1;
    return;
  }
  const double dt = a.win_dt[j], DC = a.DC;
  const int n = fqdt::num_samples(a.N, dt, DC);
  int k = (int)(a.r_fraction * (double)n);
  k = k < 0 ? 0 : (k > n - 1 ? n - 1 : k);
  double t = 0;
  int interval = 0;
  for (int i = 0; i <= k; i++)
  {
    t = FQ_ADD(t, DC);
    if (t > FQ_MUL(dt, (double)(interval + 1))) interval = interval + 1 < a.N - 1 ? interval + 1 : a.N - 1;
  }
  const double tau = FQ_SUB(t, FQ_MUL((double)interval, dt));
  double x[12], o[12];
  for (int i = 0; i < 12; i++) x[i] = a.coeffs[((size_t)j * a.N + interval) * 12 + i];
  fqdt::eval_sample(x, tau, o);
  if (k == n - 1)
    for (int i = 3; i < 12; i++) o[i] = 0.0;                       // :165-167: the last sample is at rest
  for (int i = 0; i < 9; i++) a.x0_safe[j * 9 + i] = o[i];
  a.n_samples[j] = n; a.k_safe[j] = k;
}

// result records (fq_pair_result of include/faster_b200.h; 16 doubles each)

__global__ void fq_pair_final_kernel(const FqPairFinalArgs a)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n_prob) return;
  fq_pair_result r;
  const int ww = a.win_idx_w[j], ws = a.win_idx_s[j];
  r.whole_dt_index = ww >= 0 ? ww / a.n_sig_w : -1;
  r.whole_sigma_index = ww >= 0 ? ww % a.n_sig_w : -1;
  r.safe_dt_index = ws >= 0 ? ws / a.n_sig_s : -1;
  r.safe_sigma_index = ws >= 0 ? ws % a.n_sig_s : -1;
  r.whole_cost = a.win_cost_w[j]; r.safe_cost = a.win_cost_s[j];
  r.whole_dt = a.win_dt_w[j]; r.safe_dt = a.win_dt_s[j];
  r.whole_dt_base = a.dt_base_w[j]; r.safe_dt_base = a.dt_base_s[j];
  r.n_samples_whole = a.n_samples[j]; r.k_safe = a.k_safe[j];
  for (int i = 0; i < 9; i++) r.R[i] = a.x0_safe[j * 9 + i];
  a.out[j] = r;
}
}  // namespace fqp

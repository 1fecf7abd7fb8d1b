// TEST INFRASTRUCTURE ONLY -- the small kernels of the chained replan (faster_b200/csrc/fq_pair.cuh: time-allocation base, grid
// expansion, genNewTraj's selection per corridor, the sample R between the two sweeps, the result records) compiled for the host
// and run under the block emulation of simt_emu/simt_emu.h, block by block with the launch shapes of fq_kernels.cu.  Together
// with kernel_emu.cpp (the solve kernel) tests/test_kernel_emu_cpu.py replays fq_replan_pairs_dev's whole submission without a
// GPU.  Not a CPU path of the product.
#define FQ_EMULATE_ON_HOST 1
#define __shared__ static          // these kernels keep their block-wide scalars in static __shared__ variables
#include "simt_emu/cuda_runtime.h"

#include "fq_kernels.cuh"
#include "fq_pair.cuh"

namespace
{
template <class F>
struct Thunk
{
  F f;
  static void call(void* p) { static_cast<Thunk*>(p)->f(); }
};
// launch<<<grid, block>>>: the blocks one after the other
template <class F>
void launch(unsigned grid, unsigned block, F f)
{
  Thunk<F> t{ f };
  for (unsigned b = 0; b < grid; b++) simt::run_block(dim3(grid), dim3(block), uint3{ b, 0, 0 }, Thunk<F>::call, &t, 256 * 1024);
}
}  // namespace

extern "C" {
void emu_dtbase(int n_prob, int N, double DC, const double* x0, const double* xf, const double* lim, double* dt_base)
{ // fq_launch_dtbase
  launch((4 * n_prob + 127) / 128, 128, [=] { fqp::fq_dtbase_kernel(n_prob, N, DC, x0, xf, lim, dt_base); });
}
void emu_expand_grid(int n_prob, int N, int n_fac, int n_sig, const double* factors, const uint8_t* sig_list, const double* dt_base,
                     double* dt, uint8_t* sigma, int* cand_ofs)
{ // fq_launch_expand_grid
  long long blocks = ((long long)n_prob * n_fac * n_sig + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch((unsigned)blocks, 256, [=] { fqp::fq_expand_grid_kernel(n_prob, N, n_fac, n_sig, factors, sig_list, dt_base, dt, sigma, cand_ofs); });
}
void emu_select_multi(int n_prob, int N, int n_sig, const int* cand_ofs, const double* dt, const uint8_t* sigma, const uint8_t* feasible,
                      const double* cost, int* win_idx, double* win_cost, double* win_dt, uint8_t* win_sigma, int* win_ofs)
{ // fq_launch_select_multi
  FqSelectMultiArgs a;
  a.n_prob = n_prob; a.N = N; a.n_sig = n_sig; a.cand_ofs = cand_ofs; a.dt = dt; a.sigma = sigma; a.feasible = feasible; a.cost = cost;
  a.win_idx = win_idx; a.win_cost = win_cost; a.win_dt = win_dt; a.win_sigma = win_sigma; a.win_ofs = win_ofs;
  launch((unsigned)n_prob, 128, [=] { fqp::fq_select_multi_kernel(a); });
}
void emu_pair_mid(int n_prob, int N, double DC, double r_fraction, const double* coeffs, const double* win_dt, const int* win_idx,
                  double* x0_safe, int* n_samples, int* k_safe)
{ // fq_launch_pair_mid
  FqPairMidArgs a;
  a.n_prob = n_prob; a.N = N; a.DC = DC; a.r_fraction = r_fraction; a.coeffs = coeffs; a.win_dt = win_dt; a.win_idx = win_idx;
  a.x0_safe = x0_safe; a.n_samples = n_samples; a.k_safe = k_safe;
  launch((unsigned)(n_prob + 63) / 64, 64, [=] { fqp::fq_pair_mid_kernel(a); });
}
void emu_pair_final(int n_prob, int n_sig_w, int n_sig_s, const int* win_idx_w, const int* win_idx_s, const int* n_samples, const int* k_safe,
                    const double* win_cost_w, const double* win_cost_s, const double* win_dt_w, const double* win_dt_s,
                    const double* dt_base_w, const double* dt_base_s, const double* x0_safe, fq_pair_result* out)
{ // fq_launch_pair_final
  FqPairFinalArgs a;
  a.n_prob = n_prob; a.n_sig_w = n_sig_w; a.n_sig_s = n_sig_s; a.win_idx_w = win_idx_w; a.win_idx_s = win_idx_s; a.n_samples = n_samples;
  a.k_safe = k_safe; a.win_cost_w = win_cost_w; a.win_cost_s = win_cost_s; a.win_dt_w = win_dt_w; a.win_dt_s = win_dt_s;
  a.dt_base_w = dt_base_w; a.dt_base_s = dt_base_s; a.x0_safe = x0_safe; a.out = out;
  launch((unsigned)(n_prob + 127) / 128, 128, [=] { fqp::fq_pair_final_kernel(a); });
}
}

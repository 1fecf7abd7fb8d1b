// Device-side interface of the batched corridor-QP solver (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/faster_b200.h"

#define FQ_WARPS_PER_CTA 4
#ifndef FQ_MIN_CTAS_PER_SM
#define FQ_MIN_CTAS_PER_SM 4
#endif
#ifndef FQ_MIN_CTAS_WHOLE
#define FQ_MIN_CTAS_WHOLE 4   // whole-mode kernels for N <= 10; 5 (96 registers, 20 B of spills) measured -3.5 %
#endif
#ifndef FQ_BACKSUB_PRESCALED
#define FQ_BACKSUB_PRESCALED 1   // measured +1.0 % (tools/ab.sh, same box); 0 keeps the division on the chain
#endif
#ifndef FQ_PACKED_R
#define FQ_PACKED_R 1         // 1: the triangular factor R is stored packed (column k at k(k+1)/2), half the bytes
                              // (2: only in the safe-mode kernels, whose shared memory needs it: measured equal, 54.7 vs 54.8 M);
                              // measured alone -2.4 % (index arithmetic), but it is what makes room for FQ_WARP_ADOPT
#endif
#ifndef FQ_WARP_ADOPT
#define FQ_WARP_ADOPT 1       // 1: every warp adopts problems and stages polytope rows on its own (no block barriers):
                              // +5 % over packed R alone, +2.5 % net (tools/ab.sh, same box: 53.6 -> 54.95 M pairs/s)
#endif
#ifndef FQ_SPLIT_ROWS
#define FQ_SPLIT_ROWS 0       // 1: staged polytope rows as two arrays of 16-byte halves ([Ax Ay] and [Az b+tol]): both loads of
                              // the row scan become bank-conflict free (rows 32 bytes apart give a two-way conflict)
#endif
#ifndef FQ_CONST_PRECHECK
#define FQ_CONST_PRECHECK 1   // 1: before anything else is set up, test the three control points of segment 0 that x0 and dt fix
                              // (cp0, cp1, cp2: their plan rows are zero) against the faces of sigma[0]; a violated one
                              // refutes the candidate (what the solve would find in its first iteration, same arithmetic)
#endif
#ifndef FQ_EE_CAP_ALWAYS
#define FQ_EE_CAP_ALWAYS 0     // 1: cap the grid of an early-exit sweep even when the whole sweep fits the GPU at once (tuning)
#endif
#ifndef FQ_SCAN_UNROLL
#define FQ_SCAN_UNROLL 1       // unroll factor of the corridor-row scan (more loads in flight per warp)
#endif
#ifndef FQ_MIN_REDUX
#define FQ_MIN_REDUX 1         // 1: the ratio test's warp minimum through two integer redux.sync on the order-preserving bit
                               // pattern of the doubles (exact) instead of five dependent shuffle + min rounds.  Measured
                               // (tools/kernel_ab.py, same box, 3 rounds): whole sweep 0.3976 -> 0.3825 ms, safe 0.5338 -> 0.5110 ms
#endif
#ifndef FQ_GI_HOIST
#define FQ_GI_HOIST 0          // 1: the lane's (axis, column) of the entering row's normal computed once per candidate and the
                               // per-iteration code branch-free.  Measured +0.8 % / +1.0 % SLOWER (two more live registers): off
#endif
#ifndef FQ_ITEMS_BY_SEGMENT
#define FQ_ITEMS_BY_SEGMENT 0  // 1: the corridor item list written by two lanes per segment (even / odd faces) instead of one
                               // lane per item with a four-step search for its segment.  Measured +3.2 % / +1.3 % SLOWER: off
#endif
#ifndef FQ_MAX_N_3CTAS
#define FQ_MAX_N_3CTAS 15      // largest N whose kernels are compiled for 3 CTAs per SM (<= 168 registers); beyond: 2 CTAs.  N = 14, 15
                               // run 2 CTAs per SM anyway (shared memory), yet letting the compiler take 223 registers there measured
                               // -1 % on config 5 (45.3 vs 45.8 M candidates/s, profiles/r02ab_cfg5_ab.log): the bound stays
#endif
#ifndef FQ_LAZY_LEAVING
#define FQ_LAZY_LEAVING 0      // 1: the index of the blocking row (ballot + shuffle) is looked up only when a partial step is taken.
                               // Measured (tools/kernel_ab.py, 3 rounds, profiles/r02aa_kernel_ab.log): no change (0.3862 / 0.5063 ms
                               // against 0.3862 / 0.5069 ms): off
#endif
#ifndef FQ_SCAN_ARGMAX
#define FQ_SCAN_ARGMAX 0       // 1: the scan remembers WHICH control point of the winning (segment, face) item ranked highest (two
                               // bits of the code) instead of re-evaluating the item's four points when the entering row is decoded:
                               // same keys, same order, same choice.  Measured: whole 0.3861 (=), safe 0.5022 ms (-0.9 %); with
                               // FQ_LAZY_LEAVING whole +0.3 %, safe -1.5 %: inside the noise of the A/B, left off
#endif
#define FQ_EPS_DEP 1e-18      // squared sine below which a new normal counts as dependent on the active set
#define FQ_ZZ_FLOOR 1e-30
#define FQ_MAX_ITERS 400

struct FqKernelArgs
{
  // plan (device copies of FqPlanHost tables)
  int N, force_final, ne, nz, nw, NY, ld;
  const double* TZ;      // NY x nz
  const double* T0;      // NY x (3+ne)
  const double* FT;      // ne x 3
  // problems
  int n_prob;
  const double* x0;      // n_prob x 9
  const double* xf;      // n_prob x 9
  const double* lim;     // n_prob x 3
  const int* poly_ofs;   // n_prob+1  -> index into face_ofs
  const int* face_ofs;   // n_poly_total+1 -> row of Ab
  const double* Ab;      // rows [Ax Ay Az b]
  int max_faces;         // max total faces of one problem (sizes the shared staging area)
  int item_cap;          // corridor rows per candidate the item list must hold: N * max_faces
  // candidates
  const int* cand_ofs;   // n_prob+1
  const double* dt;
  const uint8_t* sigma;  // n_cand x N
  // outputs
  uint8_t* feasible;
  double* cost;
  double* coeffs;        // n_cand x N x 12 or nullptr
  int32_t* iters;        // or nullptr
  double row_tol;        // absolute row tolerance (default FQ_ROW_TOL; option "row_tol_1e9")
  // infeasibility-certificate memo of this launch (fq_kernels_t.cuh), or nullptr: n_prob x FQ_MEMO_NB x FQ_MEMO_BE entries
  struct FqMemoEntry* memo;
  unsigned memo_salt;    // unique per launch of a context: entries carrying another salt are stale
  // "first feasible dt wins" early exit (option "sweep_early_exit"; specialised kernel only): per problem the smallest dt
  // (bit pattern) of a candidate found feasible so far, or nullptr.  Candidates with a larger dt are not evaluated.
  unsigned long long* first_feasible;
  int sorted_dt;         // candidates of every problem are in ascending dt order (the chained replan's grids)
  int ee_width;          // candidates per time allocation (0: unknown); sizes the grid of an early-exit sweep
  // single-problem sweeps (fq_gen_new_traj*): the warp that finishes the LAST candidate runs genNewTraj's selection itself
  // and writes the winner record straight into mapped host memory -- no selection launch, no device-to-host copy.
  // sweep_done = counter of finished candidates (zeroed with the claim counters), or nullptr.
  int* sweep_done;
  int sweep_n_sigma;     // candidates per time allocation (the list is dt-major)
  int* sweep_idx;        // out (host-mapped): [0] dt index (-1 none), [1] sigma index
  double* sweep_win;     // out (host-mapped): [0] cost, [1 .. 12N] coefficients of the winner (a.coeffs must be set)
  // size-generic kernel only (fq_solve_batch_cert): per infeasible candidate, [n, violation, (row id, multiplier) x n]
  double* cert;
  int cert_stride;
};

// One shared infeasibility certificate: "at time allocation dt (bit pattern), the rows of the segments in `mask` with the
// polytopes `sigpack` (4 bits per segment) have no common point with the box rows".  w0 = salt << 32 | mask << 16 | count
// slot + 1; written last, after a fence.
struct FqMemoEntry
{
  unsigned long long w0, dt_bits, sigpack;
};
#define FQ_MEMO_NB 16         // buckets per problem (hash of dt)
#define FQ_MEMO_BE 16         // entries per bucket: one per lane of a half warp
#ifndef FQ_CERT_MEMO
#define FQ_CERT_MEMO 0        // 1 compiles the memo in.  MEASURED NEGATIVE on the bench workload (same box, tools/ab.sh, cfg4
                              // chain): compiled out 72.4 M pairs/s; compiled in but switched off 64.2 M (the bookkeeping of
                              // the active rows' segments costs 11 %); switched on 57.3 M.  With ~2400 warps working on 64
                              // corridors at once and only ~28 candidates per warp per launch, few candidates start after a
                              // useful proof exists (a scheduling model of the launch gives 9 % hits and 3 % at best), and the
                              // commonest proofs (a constant control point of segment 0 outside its polytope) already cost a
                              // single iteration.  It pays on the CPU port (8-128 threads: 142 -> 66 ms per pass), where it is on.
#endif
#define FQ_MEMO_MARGIN 1e-5   // certificates are recorded only when the violation exceeds this x (1 + sum |multipliers|)

struct FqSelectArgs
{
  int n_dt, n_sigma, N;
  const uint8_t* feasible;
  const double* cost;
  const double* coeffs;  // n_dt*n_sigma x N x 12
  int* out_idx;          // [0] dt index (-1 none), [1] sigma index
  double* out_cost;      // [0]
  double* out_coeffs;    // N x 12
};

// device-side fillX (solverGurobi.cpp:122-168) of the winner selected by fq_select_kernel
struct FqFillArgs
{
  int N, n_dt, max_samples;
  double DC;
  const double* dts;        // n_dt, device
  const int* win_idx;       // [0] winning dt index (-1: nothing to sample)
  const double* coeffs;     // N x 12 of the winner
  double* out;              // max_samples x 12
  int* n_samples;           // [0] number of samples written
};
cudaError_t fq_launch_fill(const FqFillArgs& a, cudaStream_t stream);

size_t fq_solve_smem_bytes(const FqKernelArgs& a);
// picks the size-specialised kernel when one exists (4 <= N <= 16, faces <= 2047), else the generic one;
// force_generic selects the generic kernel regardless (differential testing)
// `counters`: a.n_prob ints of device memory owned by the caller's context (per-problem claim counters of the
// persistent kernel; zeroed by the launch on `stream`)
cudaError_t fq_launch_solve(const FqKernelArgs& a, int max_cand_per_prob, cudaStream_t stream, int* counters, int sm_count,
                            bool force_generic = false, bool* used_specialised = nullptr);
bool fq_has_specialised(int N, int force_final, int max_faces);
cudaError_t fq_launch_select(const FqSelectArgs& a, cudaStream_t stream);

// ---- branch-and-bound over assignments (fq_bnb.cuh)
struct FqBnbLevel
{
  FqKernelArgs k;                  // plan + ONE problem
  int n_dt, P;
  const double* dts;               // device
  int depth, n_parents;
  const unsigned char* parents;
  const int* roots;
  unsigned char* children;
  int* n_children;
  int cap;
  unsigned long long* incumbent;
  void* leaves;                    // fqb::LeafRec[leaf_cap]: {int dt_idx, pad; double cost; uchar sigma[16]}
  int* n_leaves;
  int leaf_cap;
  int* flags;
};
size_t fq_bnb_node_bytes(int N, int force_final);     // 0 if unsupported
cudaError_t fq_launch_bnb_level(const FqBnbLevel& l, cudaStream_t stream);

// ---- chained replan (fq_pair.cuh): the small kernels between the whole and the safe sweep
struct FqSelectMultiArgs
{
  int n_prob, N, n_sig;                 // n_sig > 0: candidates are a (factor x assignment) grid, dt index = local / n_sig
  const int* cand_ofs;
  const double* dt;
  const uint8_t* sigma;
  const uint8_t* feasible;
  const double* cost;
  int* win_idx;                         // n_prob: candidate index relative to the problem's first candidate, -1 none
  double* win_cost;                     // n_prob (+inf none)
  double* win_dt;                       // n_prob (NaN none): also the dt list of the winners' re-solve
  uint8_t* win_sigma;                   // n_prob x N
  int* win_ofs;                         // n_prob + 1: 0,1,2,... (cand_ofs of the re-solve)
};

struct FqPairMidArgs
{
  int n_prob, N;
  double DC, r_fraction;
  const double* coeffs;                 // n_prob x N x 12: the whole winners
  const double* win_dt;                 // n_prob (NaN: no winner)
  const int* win_idx;
  double* x0_safe;                      // n_prob x 9 (NaN when there is no whole trajectory to branch from)
  int* n_samples;                       // n_prob
  int* k_safe;                          // n_prob
};

struct FqPairFinalArgs
{
  int n_prob, n_sig_w, n_sig_s;
  const int *win_idx_w, *win_idx_s, *n_samples, *k_safe;
  const double *win_cost_w, *win_cost_s, *win_dt_w, *win_dt_s, *dt_base_w, *dt_base_s, *x0_safe;
  fq_pair_result* out;
};

cudaError_t fq_launch_dtbase(int n_prob, int N, double DC, const double* x0, const double* xf, const double* lim,
                             double* dt_base, cudaStream_t stream);
cudaError_t fq_launch_expand_grid(int n_prob, int N, int n_fac, int n_sig, const double* factors, const uint8_t* sig_list,
                                  const double* dt_base, double* dt, uint8_t* sigma, int* cand_ofs, cudaStream_t stream);
cudaError_t fq_launch_select_multi(const FqSelectMultiArgs& a, cudaStream_t stream);
cudaError_t fq_launch_pair_mid(const FqPairMidArgs& a, cudaStream_t stream);
cudaError_t fq_launch_pair_final(const FqPairFinalArgs& a, cudaStream_t stream);

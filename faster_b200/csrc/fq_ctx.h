// Internal (not installed): the solver context behind the opaque fq_ctx of include/faster_b200.h and the helpers shared
// by the translation units of the C ABI (fq_capi.cu, fq_pair_capi.cu, fq_multi.cu).
#pragma once
#include "../../include/faster_b200.h"
#include "fq_kernels.cuh"
#include "fq_plan.h"

#include <map>
#include <string>
#include <vector>

struct FqPlanDev
{
  FqPlanHost h;
  double *TZ = nullptr, *T0 = nullptr, *FT = nullptr;
};

struct FqArena
{ // grow-only device buffer
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n)
  {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 4 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct FqPinnedArena
{
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n)
  {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 4 + 4096;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct FqComm;                    // fq_multi.cu: NCCL communicator(s) of a context (loaded lazily with dlopen)

struct fq_ctx
{
  int device = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev_head = nullptr;
  std::map<int, FqPlanDev> plans; // key N*2+force_final
  FqArena d_in, d_out, d_bnb, d_pair, d_pair_io;
  FqPinnedArena h_in, h_out;
  std::string err;
  bool force_generic = false;
  int sm_count = 0;
  int* d_counters = nullptr;      // ring of per-problem claim counters (one slot per launch in flight)
  int counters_cap = 0;           // problems per slot
  unsigned counters_pos = 0;
  bool pending = false;           // a deferred host-pointer call is still using the arenas (settled by the next call / fq_wait)
  int throughput_slices = 0;      // option "throughput_slices": launches a large host batch is cut into (0 = default 4)
  int max_poly_faces_hint = 0;    // option "max_faces_per_polytope" (device-pointer API only)
  double row_tol = FQ_ROW_TOL;    // option "row_tol_1e9": absolute row tolerance of every solve of this context
  FqMemoEntry* d_memo = nullptr;  // ring of infeasibility-certificate memos: one slot per launch in flight (as d_counters)
  unsigned memo_salt = 0;         // launch counter stamped into the memo entries (stale entries carry another value)
  bool cert_memo = true;          // option "cert_memo"
  bool early_exit = false;        // option "sweep_early_exit": candidates that cannot win genNewTraj's selection are skipped
  bool launch_sorted_dt = false;  // set by the chained replan for its next launch: candidate lists are in ascending dt order
  int launch_ee_width = 0;        // ... and hold this many candidates per time allocation
  // single-problem sweep with the selection in the solve kernel's tail (consumed by the next launch):
  int launch_sweep_n_sigma = 0;   // > 0: requested
  int* launch_sweep_idx = nullptr;       // device-visible (host-mapped) outputs
  double* launch_sweep_win = nullptr;
  bool last_launch_tail = false;  // the last launch did run the tail selection (specialised kernel)
  unsigned long long* d_first = nullptr;   // ring (as d_counters) of per-problem "smallest feasible dt so far"
  double* cert_out = nullptr;     // fq_solve_batch_cert: device buffer the generic kernel writes certificates into
  int cert_stride = 0;
  // ---- multi-GPU (fq_multi.cu)
  FqComm* comm = nullptr;         // communicator this context belongs to (one process per GPU), or nullptr
  int rank = 0, world = 1;
  std::vector<fq_ctx*> peers;     // fq_create_multi: the per-device contexts of a single-process group (this = peers[0]'s owner)
  bool is_group = false;
};
static const int kFqCounterSlots = 64;
static const int kFqMemoProbs = 256;      // launches with more problems run without the certificate memo

// ---- helpers implemented in fq_capi.cu
int fq_fail(fq_ctx* c, int code, const std::string& msg);
int fq_cuda_fail(fq_ctx* c, cudaError_t e, const char* what);
#define FQ_CUDA(call)                                             \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return fq_cuda_fail(ctx, e__, #call); \
  } while (0)
int fq_get_plan(fq_ctx* ctx, int N, int force_final, FqPlanDev** out);
void fq_fill_plan_args(const FqPlanDev& pd, FqKernelArgs* a);
// a deferred call owns the context's arenas until it has drained: every entry point that reuses them settles first
int fq_settle(fq_ctx* ctx);
// common launch of the batch solve: every pointer is a device pointer
int fq_launch_solve_ctx(fq_ctx* ctx, int N, int force_final, int n_prob, const double* d_x0, const double* d_xf,
                        const double* d_lim, const int* d_poly_ofs, const int* d_face_ofs, const double* d_Ab,
                        const int* d_cand_ofs, int max_cand, int max_faces, int max_poly_faces, const double* d_dt,
                        const uint8_t* d_sigma, uint8_t* d_feasible, double* d_cost, double* d_coeffs, int32_t* d_iters,
                        cudaStream_t stream);
inline size_t fq_align16(size_t x) { return (x + 15) & ~(size_t)15; }
// fq_multi.cu
void fq_comm_release(fq_ctx* ctx);

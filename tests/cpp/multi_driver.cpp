// Drives several GPUs through the C ABI from ONE process, the way the reference's single planner process would
// (faster.hpp:74-75 keeps its solver objects in one process): fq_create_multi builds the group (NCCL communicator inside),
// fq_replan_pairs spreads the corridors of a batch over its devices and returns every corridor's result record, the same
// bits as a single-GPU context.  Also exercises SolverGurobi::setDevices.  usage: multi_driver <n_gpus>
#include "solverGurobi.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static double frand(unsigned& s) { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / 16777216.0; }

int main(int argc, char** argv)
{
  const int n_gpus = argc > 1 ? std::atoi(argv[1]) : 2;
  const int P = 7, N = 8, NF = 6, NS = 1;                 // 7 corridors (uneven shards), no polytopes: box rows only
  std::vector<double> x0(9 * P, 0.0), xfw(9 * P, 0.0), xfs(9 * P, 0.0), lim(3 * P);
  unsigned seed = 12345;
  for (int j = 0; j < P; j++)
  {
    for (int i = 0; i < 3; i++)
    {
      x0[9 * j + i] = 4 * frand(seed) - 2; x0[9 * j + 3 + i] = 2 * frand(seed) - 1;
      xfw[9 * j + i] = x0[9 * j + i] + 6 * frand(seed) - 3; xfs[9 * j + i] = x0[9 * j + i] + 6 * frand(seed) - 3;
    }
    lim[3 * j] = 5; lim[3 * j + 1] = 5; lim[3 * j + 2] = 8;
  }
  std::vector<int> zero(P + 1, 0), fo(1, 0);
  std::vector<double> Ab(4, 0.0), fac(NF);
  for (int f = 0; f < NF; f++) fac[f] = 1.0 + f;
  std::vector<uint8_t> sig(N, 0);
  fq_pair_args a;
  std::memset(&a, 0, sizeof(a));
  a.n_prob = P; a.N_whole = N; a.N_safe = N; a.DC = 0.01; a.r_fraction = 0.5;
  a.x0 = x0.data(); a.xf_whole = xfw.data(); a.xf_safe = xfs.data(); a.lim = lim.data();
  a.poly_ofs_whole = zero.data(); a.face_ofs_whole = fo.data(); a.Ab_whole = Ab.data();
  a.poly_ofs_safe = zero.data(); a.face_ofs_safe = fo.data(); a.Ab_safe = Ab.data();
  a.n_fac_whole = NF; a.factors_whole = fac.data(); a.n_sig_whole = NS; a.sigmas_whole = sig.data();
  a.n_fac_safe = NF; a.factors_safe = fac.data(); a.n_sig_safe = NS; a.sigmas_safe = sig.data();
  std::vector<fq_pair_result> r1(P), rn(P);
  std::vector<uint8_t> f1(P * NF), fn(P * NF, 7);
  std::vector<double> c1(P * NF), cn(P * NF);
  fq_ctx* one = nullptr;
  if (fq_create(&one, 0) != 0) { std::printf("fq_create: %s\n", fq_last_error(nullptr)); return 2; }
  a.results = r1.data(); a.feasible_whole = f1.data(); a.cost_whole = c1.data();
  if (fq_replan_pairs(one, &a) != 0) { std::printf("single: %s\n", fq_last_error(one)); return 3; }
  fq_destroy(one);
  SolverGurobi sg;
  std::vector<int> devs;
  for (int i = 0; i < n_gpus; i++) devs.push_back(i);
  sg.setDevices(devs);
  fq_ctx* g = sg.context();
  if (!g) { std::printf("fq_create_multi: %s\n", fq_last_error(nullptr)); return 4; }
  int rank = -1, world = -1, ver = 0;
  fq_comm_info(g, &rank, &world, &ver);
  // the class switches the early exit on (only winners matter to genNewTraj); this comparison wants every candidate
  if (fq_set_option(g, "sweep_early_exit", 0) != 0) { std::printf("option: %s\n", fq_last_error(g)); return 6; }
  a.results = rn.data(); a.feasible_whole = fn.data(); a.cost_whole = cn.data();
  if (fq_replan_pairs(g, &a) != 0) { std::printf("group: %s\n", fq_last_error(g)); return 5; }
  int bad = world != n_gpus;
  bad += std::memcmp(r1.data(), rn.data(), sizeof(fq_pair_result) * P) != 0;
  bad += std::memcmp(f1.data(), fn.data(), f1.size()) != 0;
  bad += std::memcmp(c1.data(), cn.data(), sizeof(double) * c1.size()) != 0;
  int solved = 0;
  for (int j = 0; j < P; j++) solved += r1[j].whole_dt_index >= 0;
  std::printf("{\"n_gpus\": %d, \"world\": %d, \"nccl\": %d, \"corridors\": %d, \"whole_solved\": %d, \"mismatch\": %d}\n", n_gpus, world, ver, P,
              solved, bad);
  return bad ? 1 : 0;
}

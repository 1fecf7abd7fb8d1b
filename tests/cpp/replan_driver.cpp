// A miniature of Faster::replan()'s use of the two solver instances (faster.cpp:52-71 set-up, :406-430 whole,
// :475,:521-537 safe, :582-588 factor-window update), driven over several consecutive replans along a polyline through
// an obstacle cloud.  Input generators are the product's own host code (fq_ellipsoid_decomp); planner heuristics that
// are out of scope (findIndexH/R, JPS) are replaced by fixed rules stated below so that the Python side of the test can
// replay the identical sequence with the oracle.
//
// stdin: N DC v a j gamma  n_vert  verts(n_vert x 3)  n_obs  obs(n_obs x 3)  n_replans
// rules: whole path = next <=3 segments from the current start A; E = last vertex of that path (rest);
//        R = whole sample at 60 % of the horizon; safe path = [R, following vertices of the whole path];
//        next A = whole sample at 30 % of the horizon (vel/accel kept), path restarts at the segment A lies on.
#include "solverGurobi.hpp"
#include <cstdio>
#include <iostream>

static std::vector<LinearConstraint3D> decompose(const std::vector<Eigen::Vector3d>& path, const std::vector<double>& obs)
{
  std::vector<double> p;
  for (auto& v : path) { p.push_back(v.x()); p.push_back(v.y()); p.push_back(v.z()); }
  const int n_seg = (int)path.size() - 1;
  std::vector<int> ofs(n_seg + 1);
  std::vector<double> Ab(4 * 1024);
  const double bbox[3] = { 2, 2, 1 };
  int rows = fq_ellipsoid_decomp(p.data(), n_seg, obs.data(), (int)obs.size() / 3, bbox, 0.42, 0.0, ofs.data(), Ab.data(), 1024);
  std::vector<LinearConstraint3D> out;
  if (rows < 0) return out;
  for (int s = 0; s < n_seg; s++)
  {
    FqMatX3 A; FqVecX b;
    const int F = ofs[s + 1] - ofs[s];
    A.resize(F, 3); b.resize(F);
    for (int f = 0; f < F; f++)
    {
      const double* r = &Ab[4 * (ofs[s] + f)];
      A(f, 0) = r[0]; A(f, 1) = r[1]; A(f, 2) = r[2]; b(f) = r[3];
    }
    out.push_back(LinearConstraint3D(A, b));
  }
  return out;
}

int main()
{
  int N, n_vert, n_obs, n_rep;
  double DC, lim[3], gamma;
  std::cin >> N >> DC >> lim[0] >> lim[1] >> lim[2] >> gamma >> n_vert;
  std::vector<Eigen::Vector3d> verts(n_vert);
  for (auto& v : verts) std::cin >> v.x() >> v.y() >> v.z();
  std::cin >> n_obs;
  std::vector<double> obs(3 * n_obs);
  for (auto& o : obs) std::cin >> o;
  std::cin >> n_rep;

  SolverGurobi sg_whole_, sg_safe_;                                         // faster.hpp:74-75
  SolverGurobi* both[2] = { &sg_whole_, &sg_safe_ };
  for (int k = 0; k < 2; k++)
  { // faster.cpp:52-71
    both[k]->setN(N); both[k]->createVars(); both[k]->setDC(DC); both[k]->setBounds(lim);
    both[k]->setForceFinalConstraint(k == 0); both[k]->setFactorInitialAndFinalAndIncrement(1, 10, 1.0);
    both[k]->setVerbose(0); both[k]->setThreads(0); both[k]->setWMax(4.0);
  }
  state A;
  A.setPos(verts[0].x(), verts[0].y(), verts[0].z());
  int seg0 = 0;   // index of the segment A lies on
  std::printf("[");
  for (int rep = 0; rep < n_rep; rep++)
  {
    sg_whole_.ResetToNormalState(); sg_safe_.ResetToNormalState();          // :306-307
    std::vector<Eigen::Vector3d> wpath = { A.pos };
    for (int k = seg0 + 1; k < n_vert && (int)wpath.size() < 4; k++) wpath.push_back(verts[k]);
    if (wpath.size() < 2) break;
    state E;
    E.setPos(wpath.back().x(), wpath.back().y(), wpath.back().z());
    sg_whole_.setX0(A); sg_whole_.setXf(E); sg_whole_.setPolytopes(decompose(wpath, obs));   // :406-408
    const bool ok_w = sg_whole_.genNewTraj();                                // :418
    std::printf("%s{\"rep\": %d, \"whole\": {\"solved\": %d, \"factor\": %.17g, \"dt\": %.17g, \"cost\": %.17g, \"trials\": %d, \"n_poly\": %zu}",
                rep ? ", " : "", rep, ok_w ? 1 : 0, sg_whole_.factor_that_worked_, sg_whole_.dt_, ok_w ? sg_whole_.getCost() : -1.0,
                sg_whole_.trials_, wpath.size() - 1);
    if (!ok_w) { std::printf("}"); break; }
    sg_whole_.fillX();                                                       // :427
    const std::vector<state> Xw = sg_whole_.X_temp_;
    const state R = Xw[(size_t)(0.6 * Xw.size())];                           // stand-in for findIndexR (:474-475)
    // segment of the whole path R lies closest to (by the knot time of the sample)
    const int kR = (int)(0.6 * Xw.size());
    int segR = std::min((int)(((kR + 1) * DC) / (sg_whole_.dt_ * N / (double)(wpath.size() - 1))), (int)wpath.size() - 2);
    std::vector<Eigen::Vector3d> spath = { R.pos };
    for (size_t k = segR + 1; k < wpath.size(); k++) spath.push_back(wpath[k]);
    state M;
    M.setPos(spath.back().x(), spath.back().y(), spath.back().z());
    state Rs = R;
    sg_safe_.setX0(Rs); sg_safe_.setXf(M); sg_safe_.setPolytopes(decompose(spath, obs));     // :521-523
    sg_safe_.setForceFinalConstraint(false);                                 // :524
    const bool ok_s = sg_safe_.genNewTraj();                                 // :527
    std::printf(", \"safe\": {\"solved\": %d, \"factor\": %.17g, \"dt\": %.17g, \"cost\": %.17g, \"trials\": %d, \"n_poly\": %zu, \"x0\": [%.17g, %.17g, %.17g]}}",
                ok_s ? 1 : 0, sg_safe_.factor_that_worked_, sg_safe_.dt_, ok_s ? sg_safe_.getCost() : -1.0, sg_safe_.trials_,
                spath.size() - 1, R.pos.x(), R.pos.y(), R.pos.z());
    if (!ok_s) break;
    sg_safe_.fillX();                                                        // :536
    // time allocation windows for the next replan (:582-588)
    sg_whole_.setFactorInitialAndFinalAndIncrement(std::max(sg_whole_.factor_that_worked_ - gamma, 1.0),
                                                   sg_whole_.factor_that_worked_ + gamma, 1.0);
    sg_safe_.setFactorInitialAndFinalAndIncrement(std::max(sg_safe_.factor_that_worked_ - gamma, 1.0),
                                                  sg_safe_.factor_that_worked_ + gamma, 1.0);
    // next start: whole sample at 30 % of the horizon
    const int kA = (int)(0.3 * Xw.size());
    A = Xw[kA];
    const int segA = std::min((int)(((kA + 1) * DC) / (sg_whole_.dt_ * N / (double)(wpath.size() - 1))), (int)wpath.size() - 2);
    seg0 += segA;
  }
  std::printf("]\n");
  return 0;
}

// Drives the drop-in SolverGurobi exactly as Faster::replan() does (faster.cpp:52-71, 406-430, 521-537) on a problem
// read from stdin, prints what replan() would read back.  Input (whitespace separated):
//   N force_final DC v a j f_init f_final f_inc  x0[9] xf[9]  P  then per polytope: F, F rows "ax ay az b"
#include "solverGurobi.hpp"
#include <cstdio>
#include <iostream>

int main()
{
  int N, ff, P;
  double DC, lim[3], fi, fl, finc;
  std::cin >> N >> ff >> DC >> lim[0] >> lim[1] >> lim[2] >> fi >> fl >> finc;
  state A, E;
  double v[18];
  for (int i = 0; i < 18; i++) std::cin >> v[i];
  A.setPos(v[0], v[1], v[2]); A.setVel(v[3], v[4], v[5]); A.setAccel(v[6], v[7], v[8]);
  E.setPos(v[9], v[10], v[11]); E.setVel(v[12], v[13], v[14]); E.setAccel(v[15], v[16], v[17]);
  std::cin >> P;
  std::vector<LinearConstraint3D> polys;
  for (int p = 0; p < P; p++)
  {
    int F;
    std::cin >> F;
    FqMatX3 Am; FqVecX bm;
    Am.resize(F, 3); bm.resize(F);
    for (int f = 0; f < F; f++) std::cin >> Am(f, 0) >> Am(f, 1) >> Am(f, 2) >> bm(f);
    polys.push_back(LinearConstraint3D(Am, bm));
  }
  int mode = -1;
  if (std::cin >> mode) {}                       // optional trailing token: SolverGurobi::AssignmentMode
  SolverGurobi sg;
  if (mode >= 0) sg.setAssignmentMode((SolverGurobi::AssignmentMode)mode);
  sg.setN(N); sg.createVars(); sg.setDC(DC); sg.setBounds(lim); sg.setForceFinalConstraint(ff != 0);
  sg.setFactorInitialAndFinalAndIncrement(fi, fl, finc); sg.setVerbose(0); sg.setThreads(0); sg.setWMax(4.0);
  sg.ResetToNormalState();
  sg.setX0(A); sg.setXf(E); sg.setPolytopes(polys);
  bool solved = sg.genNewTraj();
  std::printf("{\"mode\": %d, \"bnb_nodes\": %ld, \"exact\": %d, \"solved\": %d, \"trials\": %d, \"dt\": %.17g, \"factor\": %.17g, \"cost\": %.17g, \"n_samples\": %zu, \"runtime_ms\": %.3f",
              mode, sg.getBnbNodes(), sg.lastSweepExact() ? 1 : 0, solved ? 1 : 0, sg.trials_, sg.dt_, sg.factor_that_worked_,
              solved ? sg.getCost() : -1.0, sg.X_temp_.size(), sg.runtime_ms_);
  if (solved)
  {
    sg.fillX();
    std::printf(", \"coeffs\": [");
    for (size_t i = 0; i < sg.getCoeffs().size(); i++) std::printf("%s%.17g", i ? ", " : "", sg.getCoeffs()[i]);
    std::printf("], \"assignment\": [");
    auto s = sg.getAssignment();
    for (size_t i = 0; i < s.size(); i++) std::printf("%s%d", i ? ", " : "", s[i]);
    const state& first = sg.X_temp_.front();
    const state& mid = sg.X_temp_[sg.X_temp_.size() / 2];
    const state& last = sg.X_temp_.back();
    std::printf("], \"x_first\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"x_mid\": [%.17g, %.17g, %.17g], \"x_last\": [%.17g, %.17g, %.17g, %.17g]",
                first.pos.x(), first.pos.y(), first.pos.z(), first.vel.x(), first.vel.y(), first.vel.z(), mid.pos.x(),
                mid.pos.y(), mid.pos.z(), last.pos.x(), last.pos.y(), last.pos.z(), last.vel.x());
    // a second replan with the narrowed factor window (faster.cpp:582-588), same inputs
    double ni = std::max(sg.factor_that_worked_ - 20, 1.0), nf = sg.factor_that_worked_ + 20;
    sg.setFactorInitialAndFinalAndIncrement(ni, nf, finc);
    bool again = sg.genNewTraj();
    std::printf(", \"second_solved\": %d, \"second_factor\": %.17g", again ? 1 : 0, sg.factor_that_worked_);
  }
  std::printf("}\n");
  return 0;
}

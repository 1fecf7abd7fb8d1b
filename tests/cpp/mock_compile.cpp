// Compiles include/solverGurobi.hpp against the Eigen / DecompUtil API mocks (column-major MatDNf<3>) and checks the parts
// of the class surface that need no GPU: setPolytopes' row packing, the abort flag (solverGurobi.cpp:30-39,:445,:474),
// setDistances' signature (solverGurobi.hpp:100), resetX.
#include "solverGurobi.hpp"
#include <cstdio>
#ifdef FQ_EXPECT_REFERENCE_TYPES
// the reference's own headers are in use, not the look-alikes: their include guards are defined and `state` has the
// reference's printHorizontal (faster_types.hpp:160-164)
#if !defined(DECOMP_POLYGON_H) || !defined(DATA_TYPE_H)
#error "DecompUtil's own polyhedron.h / data_type.h were expected"
#endif
static_assert(std::is_same<decltype(&state::printHorizontal), void (state::*)()>::value, "the reference's state was expected");
#endif

struct Probe : SolverGurobi
{
  const std::vector<double>& rows() const { return Ab_; }
  const std::vector<int>& ofs() const { return face_ofs_; }
};

int main()
{
  Probe sg;
  MatDNf<3> A(2, 3);
  VecDf b(2);
  A(0, 0) = 1; A(0, 1) = 2; A(0, 2) = 3; A(1, 0) = 4; A(1, 1) = 5; A(1, 2) = 6;     // column-major storage underneath
  b(0) = 7; b(1) = 8;
  std::vector<LinearConstraint3D> polys(1, LinearConstraint3D(A, b));
  sg.setN(6); sg.createVars(); sg.setDC(0.01);
  double lim[3] = { 5, 5, 8 };
  sg.setBounds(lim);
  sg.setPolytopes(polys);
  const double want[8] = { 1, 2, 3, 7, 4, 5, 6, 8 };
  int bad = 0;
  for (int i = 0; i < 8; i++) bad += sg.rows()[i] != want[i];
  bad += sg.ofs().size() != 2 || sg.ofs()[1] != 2;
  vec_Vecf<3> samples;
  sg.setDistances(samples, std::vector<double>());
  // abort flag: StopExecution() before genNewTraj() => the factor loop never runs (solverGurobi.cpp:445), the call returns
  // false with trials_ == 0 and the flag is reset (:474) -- no GPU is touched
  sg.setFactorInitialAndFinalAndIncrement(1, 10, 1);
  sg.trials_ = 99;
  sg.StopExecution();
  const bool solved = sg.genNewTraj();
  bad += solved ? 1 : 0;
  bad += sg.trials_ != 0;
  bad += sg.cb_.should_terminate_.load() ? 1 : 0;
  sg.StopExecution(); sg.ResetToNormalState();
  bad += sg.cb_.should_terminate_.load() ? 1 : 0;
  state s;
  s.setPos(1, 2, 3);
  sg.setX0(s); sg.setXf(s);
  sg.dt_ = 0.3; sg.resetX();
  bad += sg.X_temp_.size() != (size_t)fq_num_samples(6, 0.3, 0.01) || sg.X_temp_.size() < 179;   // (int)(N dt / DC): truncation as in :382-388
  std::printf("%s\n", bad ? "FAIL" : "OK");
  return bad;
}

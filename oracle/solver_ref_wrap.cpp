// TEST INFRASTRUCTURE ONLY -- C wrapper around the REFERENCE's own SolverGurobi (faster/src/solverGurobi.cpp, compiled
// unmodified from /root/reference by oracle/Makefile into oracle/_ref/libsolver_ref.so).  Gurobi itself is closed source and
// absent: oracle/stub_gurobi is a RECORDING stand-in for its C++ API, oracle/stub_eigen one for Eigen.  So everything the
// reference does AROUND the numerical solve runs as the reference wrote it -- variables and their order (:70-84), the cost
// (:86-120), initial / final / continuity rows (:332-380, :499-524), the box rows (:390-407), the binaries and indicator
// rows over the Bezier control points (:180-291, :833-862), getDTInitial's choice of time allocation (:659-759; the
// polynomial root finder is a stand-in, see stub_eigen), the factor loop of genNewTraj (:426-477), resetX / fillX (:382-388,
// :122-168) -- and a test can read the model back row by row, or let an independent solver play Gurobi's part in optimize().
#include "solverGurobi.hpp"

#include <chrono>
#include <cstring>

void (*fq_grb_optimize_hook)(FqGrbCore*) = nullptr;

namespace
{
struct Ref : public SolverGurobi
{ // the members the wrapper reads are protected in the reference's class
  using SolverGurobi::m;
  using SolverGurobi::x;
  using SolverGurobi::b;
};

state make_state(const double* s)
{
  state st;
  st.setPos(s[0], s[1], s[2]);
  st.setVel(s[3], s[4], s[5]);
  st.setAccel(s[6], s[7], s[8]);
  return st;
}
std::vector<LinearConstraint3D> make_polys(int P, const int* face_ofs, const double* Ab)
{
  std::vector<LinearConstraint3D> out;
  for (int p = 0; p < P; p++)
  {
    const int F = face_ofs[p + 1] - face_ofs[p];
    MatDNf<3> A(F, 3);
    VecDf bb(F);
    for (int f = 0; f < F; f++)
    {
      const double* r = Ab + 4 * (size_t)(face_ofs[p] + f);
      A(f, 0) = r[0]; A(f, 1) = r[1]; A(f, 2) = r[2]; bb(f) = r[3];
    }
    out.push_back(LinearConstraint3D(A, bb));
  }
  return out;
}
// the set-up sequence of Faster::Faster (faster.cpp:52-71) for one solver object
void setup(Ref& s, int N, int force_final, const double* lim, double DC)
{
  s.setN(N);
  s.createVars();
  s.setDC(DC);
  double mx[3] = { lim[0], lim[1], lim[2] };
  s.setBounds(mx);
  s.setForceFinalConstraint(force_final != 0);
}

// ---- the solve callback an outside solver implements (Python, tests): dense description of the recorded model
typedef int (*solve_cb_t)(int n_vars, int n_rows, const double* A, const int* sense, const double* rhs, const int* ind_var, const int* ind_val,
                          const char* vtype, const double* qdiag, double* x_out, double* obj_out);
solve_cb_t g_cb = nullptr;
int g_offdiag = 0;

void dump(const FqGrbCore* c, std::vector<double>& A, std::vector<int>& sense, std::vector<double>& rhs, std::vector<int>& iv,
          std::vector<int>& ival, std::vector<char>& vt, std::vector<double>& qd, int* n_rows)
{
  const int nv = (int)c->vars.size();
  vt.resize(nv);
  for (int i = 0; i < nv; i++) vt[i] = c->vars[i].removed ? 'R' : c->vars[i].type;
  qd.assign(nv, 0.0);
  g_offdiag = 0;
  for (const auto& q : c->qobj) { if (q.i == q.j) qd[q.i] += q.c; else g_offdiag++; }
  int nr = 0;
  for (const auto& r : c->rows) if (!r.removed) nr++;
  A.assign((size_t)nr * nv, 0.0); sense.resize(nr); rhs.resize(nr); iv.resize(nr); ival.resize(nr);
  int k = 0;
  for (const auto& r : c->rows)
  {
    if (r.removed) continue;
    for (const auto& t : r.e.terms) A[(size_t)k * nv + t.first] = t.second;
    sense[k] = r.sense; rhs[k] = r.rhs; iv[k] = r.ind_var; ival[k] = r.ind_val;
    k++;
  }
  *n_rows = nr;
}
void hook(FqGrbCore* c)
{
  if (!g_cb) return;
  std::vector<double> A, rhs, qd;
  std::vector<int> sense, iv, ival;
  std::vector<char> vt;
  int nr = 0;
  dump(c, A, sense, rhs, iv, ival, vt, qd, &nr);
  const int nv = (int)c->vars.size();
  std::vector<double> xo(nv, 0.0);
  double obj = 0;
  const int st = g_cb(nv, nr, A.data(), sense.data(), rhs.data(), iv.data(), ival.data(), vt.data(), qd.data(), xo.data(), &obj);
  c->status = st;
  if (st == GRB_OPTIMAL)
  {
    for (int i = 0; i < nv; i++) c->vars[i].value = xo[i];
    c->objval = obj;
  }
}
}  // namespace

extern "C" {
// The model the reference builds for ONE trial at time allocation dt (genNewTraj's body, :445-455, after findDT).
// Outputs (row-major; n_vars = 12 N continuous + (N+1) P binaries, the reference's creation order):
//   A[n_rows x n_vars], sense ('=' '<' '>'), rhs, ind_var (binary variable of an indicator row, -1 for a plain row), ind_val,
//   vtype[n_vars] ('C' 'B'), qdiag[n_vars] (objective = sum qdiag_i x_i^2).  Returns n_rows, -1 if cap_rows is too small;
//   *n_vars_out, *offdiag_out = number of off-diagonal objective terms (0 for this model).
int solverref_model(int N, int force_final, const double* x0, const double* xf, const double* lim, double DC, double dt, int P,
                    const int* face_ofs, const double* Ab, double* A, int* sense, double* rhs, int* ind_var, int* ind_val, char* vtype,
                    double* qdiag, int cap_rows, int cap_vars, int* n_vars_out, int* offdiag_out)
{
  Ref s;
  setup(s, N, force_final, lim, DC);
  state a = make_state(x0), b = make_state(xf);
  s.setX0(a);                                                    // faster.cpp:406-408 / :521-524
  s.setXf(b);
  s.setPolytopes(make_polys(P, face_ofs, Ab));
  s.dt_ = dt;                                                    // findDT's result, given
  s.setPolytopesConstraints();                                   // :448-452
  s.setConstraintsX0();
  s.setConstraintsXf();
  s.setDynamicConstraints();
  s.setObjective();
  std::vector<double> Av, rv, qd;
  std::vector<int> sv, iv, ival;
  std::vector<char> vt;
  int nr = 0;
  dump(s.m.core.get(), Av, sv, rv, iv, ival, vt, qd, &nr);
  const int nv = (int)vt.size();
  if (n_vars_out) *n_vars_out = nv;
  if (offdiag_out) *offdiag_out = g_offdiag;
  if (nr > cap_rows || nv > cap_vars) return -1;
  std::memcpy(A, Av.data(), sizeof(double) * Av.size());
  std::memcpy(sense, sv.data(), sizeof(int) * nr); std::memcpy(rhs, rv.data(), sizeof(double) * nr);
  std::memcpy(ind_var, iv.data(), sizeof(int) * nr); std::memcpy(ind_val, ival.data(), sizeof(int) * nr);
  std::memcpy(vtype, vt.data(), nv); std::memcpy(qdiag, qd.data(), sizeof(double) * nv);
  return nr;
}

// Seconds the reference's own per-trial set-up takes (the body of genNewTraj's loop without the solve, :445-455: findDT,
// setPolytopesConstraints, setConstraintsX0/Xf, setDynamicConstraints, setObjective, resetX), repeated n_trials times on one
// solver object like consecutive factors of a sweep.  The stand-in's objects are lighter than Gurobi's, so this is a lower bound
// on what the reference spends per trial before Gurobi even starts.
double solverref_time_setup(int N, int force_final, const double* x0, const double* xf, const double* lim, double DC, int P,
                            const int* face_ofs, const double* Ab, int n_trials)
{
  Ref s;
  setup(s, N, force_final, lim, DC);
  state a = make_state(x0), b = make_state(xf);
  s.setX0(a);
  s.setXf(b);
  s.setPolytopes(make_polys(P, face_ofs, Ab));
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n_trials; i++)
  {
    s.findDT(1.0 + i);
    s.setPolytopesConstraints();
    s.setConstraintsX0();
    s.setConstraintsXf();
    s.setDynamicConstraints();
    s.setObjective();
    s.resetX();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// getDTInitial (:659-759) for (x0, xf, limits, N)
double solverref_dt_initial(int N, const double* x0, const double* xf, const double* lim)
{
  Ref s;
  setup(s, N, 1, lim, 0.01);
  state a = make_state(x0), b = make_state(xf);
  s.setX0(a);
  s.setXf(b);
  return s.getDTInitial();
}

// resetX + fillX (:382-388, :122-168) on given coefficients (N x 12, x[t][i] order of :72).  out: n x 12 (pos vel accel jerk).
int solverref_fill_x(int N, const double* coeffs, double dt, double DC, double* out, int cap)
{
  Ref s;
  double lim[3] = { 1, 1, 1 };
  setup(s, N, 1, lim, DC);
  for (int t = 0; t < N; t++)
    for (int i = 0; i < 12; i++) s.m.core->vars[(size_t)s.x[t][i].id].value = coeffs[12 * t + i];
  s.dt_ = dt;
  s.resetX();
  s.fillX();
  const int n = (int)s.X_temp_.size();
  for (int i = 0; i < n && i < cap; i++)
  {
    const state& q = s.X_temp_[i];
    double* o = out + 12 * (size_t)i;
    for (int k = 0; k < 3; k++) { o[k] = q.pos(k); o[3 + k] = q.vel(k); o[6 + k] = q.accel(k); o[9 + k] = q.jerk(k); }
  }
  return n;
}

// The reference's genNewTraj (:426-477) end to end, `cb` playing Gurobi's part in optimize().  stop_first != 0: StopExecution()
// is called before (the abort flag, :30-39).  Returns solved; out: trials_, dt_, factor_that_worked_, coefficients (N x 12),
// samples of fillX (n x 12, only when solved), number of optimize() calls.
int solverref_gen_new_traj(int N, int force_final, const double* x0, const double* xf, const double* lim, double DC, int P,
                           const int* face_ofs, const double* Ab, double f_init, double f_final, double f_inc, solve_cb_t cb,
                           int stop_first, int* trials, double* dt, double* factor, double* coeffs, double* samples, int cap,
                           int* n_samples, int* n_optimize)
{
  Ref s;
  setup(s, N, force_final, lim, DC);
  s.setFactorInitialAndFinalAndIncrement(f_init, f_final, f_inc);
  state a = make_state(x0), b = make_state(xf);
  s.setX0(a);
  s.setXf(b);
  s.setPolytopes(make_polys(P, face_ofs, Ab));
  g_cb = cb;
  fq_grb_optimize_hook = hook;
  if (stop_first) s.StopExecution();
  const bool solved = s.genNewTraj();
  fq_grb_optimize_hook = nullptr;
  g_cb = nullptr;
  *trials = s.trials_; *dt = s.dt_; *factor = s.factor_that_worked_;
  if (n_optimize) *n_optimize = s.m.core->optimizations;
  *n_samples = 0;
  if (solved)
  {
    for (int t = 0; t < N; t++)
      for (int i = 0; i < 12; i++) coeffs[12 * t + i] = s.x[t][i].get(GRB_DoubleAttr_X);
    s.fillX();                                                   // faster.cpp:427 / :536
    const int n = (int)s.X_temp_.size();
    *n_samples = n;
    for (int i = 0; i < n && i < cap; i++)
    {
      const state& q = s.X_temp_[i];
      double* o = samples + 12 * (size_t)i;
      for (int k = 0; k < 3; k++) { o[k] = q.pos(k); o[3 + k] = q.vel(k); o[6 + k] = q.accel(k); o[9 + k] = q.jerk(k); }
    }
  }
  return solved ? 1 : 0;
}
}

// TEST INFRASTRUCTURE ONLY -- a RECORDING stand-in for the Gurobi C++ API (closed source, absent from this image), exactly as
// wide as faster/src/solverGurobi.cpp uses it.  It does not solve anything: variables, linear (in)equalities, indicator
// constraints and the quadratic objective that the REFERENCE'S OWN model-building code creates are kept as plain data, so that
// the reference's solverGurobi.cpp compiles unmodified from /root/reference (oracle/Makefile -> oracle/_ref/libsolver_ref.so)
// and a test can read back the model it builds, row by row (oracle/solver_ref_wrap.cpp).  optimize() hands the recorded
// model to a hook the wrapper installs (an independent solver playing Gurobi's part), or reports "infeasible" without one.
#pragma once
#include <map>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

#define GRB_INFINITY 1e100
#define GRB_CONTINUOUS 'C'
#define GRB_BINARY 'B'
#define GRB_INTEGER 'I'
#define GRB_MINIMIZE 1
#define GRB_MAXIMIZE -1
#define GRB_LESS_EQUAL '<'
#define GRB_GREATER_EQUAL '>'
#define GRB_EQUAL '='
#define GRB_LOADED 1
#define GRB_OPTIMAL 2
#define GRB_INFEASIBLE 3
#define GRB_INF_OR_UNBD 4
#define GRB_UNBOUNDED 5
#define GRB_INTERRUPTED 11
#define GRB_NUMERIC 12
enum GRB_StringAttr { GRB_StringAttr_ModelName };
enum GRB_DoubleAttr { GRB_DoubleAttr_Runtime, GRB_DoubleAttr_X, GRB_DoubleAttr_ObjVal };
enum GRB_IntAttr { GRB_IntAttr_Status };

struct FqGrbVar { double lb, ub, value; char type; std::string name; bool removed; };
struct FqGrbLin { std::map<int, double> terms; double c = 0; };
struct FqGrbRow { FqGrbLin e; char sense; double rhs; std::string name; bool removed; int ind_var; int ind_val; };   // ind_var < 0: plain
struct FqGrbQuadTerm { int i, j; double c; };
struct FqGrbCore
{
  std::vector<FqGrbVar> vars;
  std::vector<FqGrbRow> rows;            // linear and indicator constraints in creation order
  std::vector<FqGrbQuadTerm> qobj;
  FqGrbLin lobj;
  int sense = GRB_MINIMIZE, status = GRB_LOADED, updates = 0, optimizations = 0;
  double runtime = 0, objval = 0;
  std::map<std::string, std::string> params;
  std::string name;
};
// installed by the wrapper: plays Gurobi's part in optimize() (fills vars[].value, status, objval); may stay null
extern "C" { extern void (*fq_grb_optimize_hook)(FqGrbCore*); }

class GRBException
{
public:
  GRBException(std::string m = "", int c = 0) : msg_(m), code_(c) {}
  std::string getMessage() const { return msg_; }
  int getErrorCode() const { return code_; }
private:
  std::string msg_; int code_;
};
class GRBEnv { public: GRBEnv() {} };
class GRBCallback
{
public:
  virtual ~GRBCallback() {}
  bool fq_aborted = false;
protected:
  int where = 0;
  virtual void callback() {}
  void abort() { fq_aborted = true; }
  friend class GRBModel;
};

class GRBVar
{
public:
  GRBVar() {}
  GRBVar(std::shared_ptr<FqGrbCore> c, int i) : core(c), id(i) {}
  double get(GRB_DoubleAttr) const { return core->vars[(size_t)id].value; }
  std::shared_ptr<FqGrbCore> core;
  int id = -1;
};

class GRBLinExpr
{
public:
  GRBLinExpr(double c = 0) { e.c = c; }
  GRBLinExpr(const GRBVar& v, double coeff = 1.0) : core(v.core) { e.terms[v.id] = coeff; }
  double getValue() const
  {
    double s = e.c;
    for (const auto& t : e.terms) s += t.second * core->vars[(size_t)t.first].value;
    return s;
  }
  GRBLinExpr& operator+=(const GRBLinExpr& o)
  {
    if (!core) core = o.core;
    for (const auto& t : o.e.terms) e.terms[t.first] += t.second;
    e.c += o.e.c;
    return *this;
  }
  GRBLinExpr& operator*=(double s)
  {
    for (auto& t : e.terms) t.second *= s;
    e.c *= s;
    return *this;
  }
  FqGrbLin e;
  std::shared_ptr<FqGrbCore> core;
};
inline GRBLinExpr operator+(GRBLinExpr a, const GRBLinExpr& b) { a += b; return a; }
inline GRBLinExpr operator-(const GRBLinExpr& a) { GRBLinExpr r = a; r *= -1.0; return r; }
inline GRBLinExpr operator-(GRBLinExpr a, const GRBLinExpr& b) { a += -b; return a; }
inline GRBLinExpr operator*(GRBLinExpr a, double s) { a *= s; return a; }
inline GRBLinExpr operator*(double s, GRBLinExpr a) { a *= s; return a; }
inline GRBLinExpr operator/(GRBLinExpr a, double s)
{ // every coefficient divided (how Gurobi's own classes round here is not public: tests allow a few ulps on such rows)
  for (auto& t : a.e.terms) t.second /= s;
  a.e.c /= s;
  return a;
}
inline GRBLinExpr operator*(const GRBVar& v, double s) { return GRBLinExpr(v, s); }
inline GRBLinExpr operator*(double s, const GRBVar& v) { return GRBLinExpr(v, s); }
inline GRBLinExpr operator+(const GRBVar& a, const GRBVar& b) { return GRBLinExpr(a) + GRBLinExpr(b); }
inline std::ostream& operator<<(std::ostream& os, const GRBLinExpr& x)
{
  os << x.e.c;
  for (const auto& t : x.e.terms) os << " + " << t.second << " v" << t.first;
  return os;
}

class GRBQuadExpr
{
public:
  GRBQuadExpr(double c = 0) { lin.e.c = c; }
  GRBQuadExpr(const GRBLinExpr& l) : lin(l) {}
  GRBQuadExpr& operator+=(const GRBQuadExpr& o)
  {
    q.insert(q.end(), o.q.begin(), o.q.end());
    lin += o.lin;
    return *this;
  }
  std::vector<FqGrbQuadTerm> q;
  GRBLinExpr lin;
};
inline GRBQuadExpr operator+(GRBQuadExpr a, const GRBQuadExpr& b) { a += b; return a; }
inline GRBQuadExpr operator*(const GRBLinExpr& a, const GRBLinExpr& b)
{
  GRBQuadExpr r(a.e.c * b.e.c);
  for (const auto& s : a.e.terms)
    for (const auto& t : b.e.terms) r.q.push_back({ s.first, t.first, s.second * t.second });
  GRBLinExpr la = a, lb = b;
  la.e.c = 0; lb.e.c = 0;
  r.lin += la * b.e.c;
  r.lin += lb * a.e.c;
  return r;
}
inline GRBQuadExpr operator*(double s, GRBQuadExpr a)
{
  for (auto& t : a.q) t.c *= s;
  a.lin *= s;
  return a;
}

class GRBTempConstr
{
public:
  GRBLinExpr e;      // e (sense) 0
  char sense;
};
inline GRBTempConstr operator==(const GRBLinExpr& a, const GRBLinExpr& b) { return GRBTempConstr{ a - b, GRB_EQUAL }; }
inline GRBTempConstr operator<=(const GRBLinExpr& a, const GRBLinExpr& b) { return GRBTempConstr{ a - b, GRB_LESS_EQUAL }; }
inline GRBTempConstr operator>=(const GRBLinExpr& a, const GRBLinExpr& b) { return GRBTempConstr{ a - b, GRB_GREATER_EQUAL }; }

class GRBConstr { public: int id = -1; };
class GRBGenConstr { public: int id = -1; };
class GRBQConstr { public: int id = -1; };

class GRBModel
{
public:
  explicit GRBModel(const GRBEnv&) : core(std::make_shared<FqGrbCore>()) {}
  void set(GRB_StringAttr, const std::string& v) { core->name = v; }
  void set(const std::string& k, const std::string& v) { core->params[k] = v; }
  void setCallback(GRBCallback* cb) { cb_ = cb; }
  GRBVar addVar(double lb, double ub, double, char type, std::string name)
  {
    core->vars.push_back(FqGrbVar{ lb, ub, 0.0, type, name, false });
    return GRBVar(core, (int)core->vars.size() - 1);
  }
  GRBConstr addConstr(const GRBTempConstr& t, std::string name = "")
  { // stored as  terms (sense) rhs  with the constant moved to the right-hand side
    FqGrbRow r{ t.e.e, t.sense, -t.e.e.c, name, false, -1, 0 };
    r.e.c = 0;
    core->rows.push_back(r);
    GRBConstr h; h.id = (int)core->rows.size() - 1;
    return h;
  }
  GRBGenConstr addGenConstrIndicator(GRBVar bin, int val, const GRBLinExpr& e, char sense, double rhs, std::string name = "")
  {
    FqGrbRow r{ e.e, sense, rhs - e.e.c, name, false, bin.id, val };
    r.e.c = 0;
    core->rows.push_back(r);
    GRBGenConstr h; h.id = (int)core->rows.size() - 1;
    return h;
  }
  void remove(GRBConstr c) { core->rows[(size_t)c.id].removed = true; }
  void remove(GRBGenConstr c) { core->rows[(size_t)c.id].removed = true; }
  void remove(GRBQConstr) {}
  void remove(GRBVar v) { core->vars[(size_t)v.id].removed = true; }
  void setObjective(const GRBQuadExpr& q, int sense)
  {
    core->qobj = q.q;
    core->lobj = q.lin.e;
    core->sense = sense;
  }
  void update() { core->updates++; }
  void optimize()
  {
    core->optimizations++;
    core->status = GRB_INFEASIBLE;
    if (cb_) { cb_->where = 0; cb_->callback(); if (cb_->fq_aborted) { cb_->fq_aborted = false; core->status = GRB_INTERRUPTED; return; } }
    if (fq_grb_optimize_hook) fq_grb_optimize_hook(core.get());
  }
  double get(GRB_DoubleAttr a) const { return a == GRB_DoubleAttr_Runtime ? core->runtime : core->objval; }
  int get(GRB_IntAttr) const { return core->status; }
  std::shared_ptr<FqGrbCore> core;
private:
  GRBCallback* cb_ = nullptr;
};

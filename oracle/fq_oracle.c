/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of FASTER's SolverGurobi hot path.
 *
 * PARITY UNPINNED: the arithmetic of the reference lives in Gurobi (closed source, version not pinned by
 * faster/FindGUROBI.cmake:8-16, Readme.md:43 lists 8.1/9.0/9.1; absent from /root/reference and from this
 * image) and the reference tree records no outputs for this path.  This file restates the
 * *model* the reference hands to Gurobi and solves it exactly; it is pinned against (i) the closed form of
 * config 1 (N=3, zero degrees of freedom) and (ii) an independent solver (HiGHS) run on the literal
 * full-space model of oracle/model_fullspace.py -- which in turn equals, row by row, the model the reference's own
 * solverGurobi.cpp builds (compiled over a recording Gurobi stand-in: oracle/solver_ref.py); this file's getDTInitial,
 * fillX and genNewTraj sweep are compared with the reference's own functions there too (tests/test_reference_solver_cpu.py).
 * Labelled "CPU restatement of SolverGurobi", never "Gurobi".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (faster_b200/, include/) never links, imports or calls it.
 *
 * What is restated (reference file:line):
 *   variables  x[t][0..11] = (ax ay az bx by bz cx cy cz dx dy dz)     solverGurobi.cpp:70-84
 *   cost       sum_t sum_axis (6 a)^2                                   solverGurobi.cpp:113-119, :783-788
 *   initial    d0=p0, c0=v0, 2 b0=a0                                    solverGurobi.cpp:359-380
 *   final      p(dt)=pf (if forceFinalConstraint_), v(dt)=vf, a(dt)=af  solverGurobi.cpp:332-357
 *   continuity p,v,a at interior knots                                  solverGurobi.cpp:499-524
 *   boxes      |c|<=vmax, |2b|<=amax, |6a|<=jmax at segment starts      solverGurobi.cpp:390-407
 *   corridor   binary b[t][p]; b=1 => A_p cp_k(t) <= b_p, k=0..3        solverGurobi.cpp:180-291, :833-862
 *   sweep      ascending factor, first feasible wins                    solverGurobi.cpp:426-477, :494-497
 *   dt guess   getDTInitial (float temporaries), MinPositiveElement     solverGurobi.cpp:659-759, utils.hpp:19-32
 *   sampling   resetX / fillX                                           solverGurobi.cpp:382-388, :122-168
 *
 * Method.  Initial + continuity equalities are satisfied by construction: the 12N coefficients are an
 * affine function z = zc + M u of the 3N jerks u (a_t = u_t/6 and the continuity rows propagate b,c,d).
 * Every remaining row is written over the 12N coefficients exactly as the reference writes it and mapped
 * to u-space through M.  For a fixed assignment sigma the problem is a strictly convex QP (cost u'u) solved by a
 * dense Goldfarb-Idnani dual active-set method (exact, detects infeasibility).  The MIQP optimum for one dt is
 * the minimum over sigma, found by depth-first branch-and-bound over segments with dual warm starts (every
 * sigma in P^N is reachable, no monotonicity assumption) or by enumerating a caller-supplied sigma list.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>

#define FQO_MAXN 16
#define FQO_NV (3 * FQO_MAXN)          /* jerk variables            */
#define FQO_NZ (12 * FQO_MAXN)         /* spline coefficients       */
static double g_row_tol = 1e-8;        /* row violation tolerance (row units: m, m/s, m/s2, m/s3); mirrors the
                                          product's option "row_tol_1e9" (set before a batch, read by the workers)  */
#define FQO_TOL g_row_tol
#define FQO_EPS_DEP 1e-18              /* squared sine below which a normal counts as dependent   */
#define FQO_MAX_ITER 2000

typedef struct
{
  int N, force_final, P, SF;           /* SF = total faces                                       */
  const int* face_ofs;                 /* P+1                                                    */
  const double* Ab;                    /* SF x 4 row-major [Ax Ay Az b]                          */
  double x0[9], xf[9], lim[3], dt;
  int n;                               /* 3N                                                     */
  double zc[FQO_NZ];                   /* coefficients at u = 0                                  */
  double M[FQO_NZ][FQO_NV];            /* d coefficients / d u                                   */
} fqo_model;

typedef struct
{
  double x[FQO_NV];                    /* primal (jerks)                                         */
  double J[FQO_NV][FQO_NV];            /* orthogonal; first q columns span the active normals    */
  double R[FQO_NV][FQO_NV];            /* upper triangular q x q                                 */
  double lam[FQO_NV];
  int act[FQO_NV];                     /* row ids; equalities have id < 0                        */
  int q, neq;
} fqo_state;

/* ---------------------------------------------------------------------------------------------
 * coefficient map  z = zc + M u      (initial + continuity rows hold by construction)
 * ------------------------------------------------------------------------------------------- */
static void model_build_map(fqo_model* m)
{
  const int N = m->N, n = 3 * N;
  const double dt = m->dt;
  for (int i = 0; i < 12 * N; i++)
  {
    m->zc[i] = 0;
    memset(m->M[i], 0, sizeof(double) * n);
  }
  for (int ax = 0; ax < 3; ax++)
  {
    /* segment 0: d = p0, c = v0, b = a0/2  (solverGurobi.cpp:369-379) */
    m->zc[9 + ax] = m->x0[ax];
    m->zc[6 + ax] = m->x0[3 + ax];
    m->zc[3 + ax] = m->x0[6 + ax] / 2.0;
    for (int t = 0; t < N; t++)
    {
      double* Ma = m->M[12 * t + 0 + ax];
      Ma[3 * t + ax] = 1.0 / 6.0;      /* jerk = 6 a  (solverGurobi.cpp:786) */
      if (t + 1 < N)
      {
        const int a = 12 * t + 0 + ax, b = 12 * t + 3 + ax, c = 12 * t + 6 + ax, d = 12 * t + 9 + ax;
        const int b1 = b + 12, c1 = c + 12, d1 = d + 12;
        /* continuity (solverGurobi.cpp:513-521): d' = p(dt); c' = v(dt); 2 b' = acc(dt) */
        m->zc[d1] = m->zc[a] * dt * dt * dt + m->zc[b] * dt * dt + m->zc[c] * dt + m->zc[d];
        m->zc[c1] = 3 * m->zc[a] * dt * dt + 2 * m->zc[b] * dt + m->zc[c];
        m->zc[b1] = (6 * m->zc[a] * dt + 2 * m->zc[b]) / 2.0;
        for (int j = 0; j < n; j++)
        {
          m->M[d1][j] = m->M[a][j] * dt * dt * dt + m->M[b][j] * dt * dt + m->M[c][j] * dt + m->M[d][j];
          m->M[c1][j] = 3 * m->M[a][j] * dt * dt + 2 * m->M[b][j] * dt + m->M[c][j];
          m->M[b1][j] = (6 * m->M[a][j] * dt + 2 * m->M[b][j]) / 2.0;
        }
      }
    }
  }
}

/* sparse functional over the 12N coefficients */
typedef struct
{
  int nnz;
  int idx[12];
  double val[12];
} fqo_row;

static void row_pos(fqo_row* r, int t, double tau, int ax, double w)
{ /* solverGurobi.cpp:761-767 */
  int k = r->nnz;
  r->idx[k] = 12 * t + 0 + ax; r->val[k++] = w * tau * tau * tau;
  r->idx[k] = 12 * t + 3 + ax; r->val[k++] = w * tau * tau;
  r->idx[k] = 12 * t + 6 + ax; r->val[k++] = w * tau;
  r->idx[k] = 12 * t + 9 + ax; r->val[k++] = w;
  r->nnz = k;
}
static void row_vel(fqo_row* r, int t, double tau, int ax, double w)
{ /* solverGurobi.cpp:769-774 */
  int k = r->nnz;
  r->idx[k] = 12 * t + 0 + ax; r->val[k++] = w * 3 * tau * tau;
  r->idx[k] = 12 * t + 3 + ax; r->val[k++] = w * 2 * tau;
  r->idx[k] = 12 * t + 6 + ax; r->val[k++] = w;
  r->nnz = k;
}
static void row_acc(fqo_row* r, int t, double tau, int ax, double w)
{ /* solverGurobi.cpp:776-781 */
  int k = r->nnz;
  r->idx[k] = 12 * t + 0 + ax; r->val[k++] = w * 6 * tau;
  r->idx[k] = 12 * t + 3 + ax; r->val[k++] = w * 2;
  r->nnz = k;
}
static void row_jerk(fqo_row* r, int t, int ax, double w)
{ /* solverGurobi.cpp:783-788 */
  int k = r->nnz;
  r->idx[k] = 12 * t + 0 + ax; r->val[k++] = w * 6;
  r->nnz = k;
}
static void row_cp(fqo_row* r, int t, int k, double dt, int ax, double w)
{ /* control points, solverGurobi.cpp:812-862 */
  if (k == 0) { row_pos(r, t, 0.0, ax, w); return; }
  if (k == 3) { row_pos(r, t, dt, ax, w); return; }
  int i = r->nnz;
  if (k == 1)
  { /* (Cn + 3 Dn)/3 */
    r->idx[i] = 12 * t + 6 + ax; r->val[i++] = w * dt / 3.0;
    r->idx[i] = 12 * t + 9 + ax; r->val[i++] = w * 3.0 / 3.0;
  }
  else
  { /* (Bn + 2 Cn + 3 Dn)/3 */
    r->idx[i] = 12 * t + 3 + ax; r->val[i++] = w * dt * dt / 3.0;
    r->idx[i] = 12 * t + 6 + ax; r->val[i++] = w * 2.0 * dt / 3.0;
    r->idx[i] = 12 * t + 9 + ax; r->val[i++] = w * 3.0 / 3.0;
  }
  r->nnz = i;
}

/* Row ids.  Box rows: id = ((t*3+ax)*3+type)*2+sign, type 0 v, 1 a, 2 j (setMaxConstraints order).
 * Corridor rows: id = NB + ((t*4+k)*SF + gf), gf = global face index.  Form: row.z <= rhs. */
static int n_box(const fqo_model* m) { return 18 * m->N; }

static double build_row(const fqo_model* m, int id, fqo_row* r)
{
  r->nnz = 0;
  const int NB = n_box(m);
  if (id < NB)
  {
    int sign = id & 1, type = (id >> 1) % 3, tax = (id >> 1) / 3, ax = tax % 3, t = tax / 3;
    double w = sign ? -1.0 : 1.0;
    if (type == 0) row_vel(r, t, 0.0, ax, w);
    else if (type == 1) row_acc(r, t, 0.0, ax, w);
    else row_jerk(r, t, ax, w);
    return m->lim[type];
  }
  int c = id - NB, gf = c % m->SF, tk = c / m->SF, k = tk & 3, t = tk >> 2;
  const double* ab = m->Ab + 4 * gf;
  for (int ax = 0; ax < 3; ax++) row_cp(r, t, k, m->dt, ax, ab[ax]);
  return ab[3];
}

static void row_to_u(const fqo_model* m, const fqo_row* r, double* g, double* c0)
{ /* g = M' row,  c0 = row . zc */
  const int n = m->n;
  double c = 0;
  for (int j = 0; j < n; j++) g[j] = 0;
  for (int k = 0; k < r->nnz; k++)
  {
    const double v = r->val[k];
    const double* Mr = m->M[r->idx[k]];
    c += v * m->zc[r->idx[k]];
    for (int j = 0; j < n; j++) g[j] += v * Mr[j];
  }
  *c0 = c;
}

/* ---------------------------------------------------------------------------------------------
 * Goldfarb-Idnani core (G = I, cost 1/2 u'u; rows g.u <= h)
 * ------------------------------------------------------------------------------------------- */
static void gi_init(fqo_state* s, int n)
{
  for (int i = 0; i < n; i++)
  {
    memset(s->J[i], 0, sizeof(double) * n);
    memset(s->R[i], 0, sizeof(double) * n);
    s->J[i][i] = 1.0;
    s->x[i] = 0; s->lam[i] = 0; s->act[i] = 0;
  }
  s->q = 0; s->neq = 0;
}

static void gi_add(fqo_state* s, int n, double* d, int id, double lam)
{ /* zero d[q+1..n-1] with column rotations of J, append d[0..q] as a column of R */
  const int q = s->q;
  for (int j = n - 1; j > q; j--)
  {
    double a = d[j - 1], b = d[j];
    if (b == 0.0) continue;
    double h = hypot(a, b), c = a / h, sn = b / h;
    d[j - 1] = h; d[j] = 0.0;
    for (int i = 0; i < n; i++)
    {
      double u = s->J[i][j - 1], v = s->J[i][j];
      s->J[i][j - 1] = c * u + sn * v;
      s->J[i][j] = -sn * u + c * v;
    }
  }
  for (int i = 0; i <= q; i++) s->R[i][q] = d[i];
  s->act[q] = id; s->lam[q] = lam; s->q = q + 1;
}

static void gi_drop(fqo_state* s, int n, int l)
{
  const int q = s->q;
  for (int j = l; j < q - 1; j++)
  {
    for (int i = 0; i <= j + 1; i++) s->R[i][j] = s->R[i][j + 1];
    s->act[j] = s->act[j + 1]; s->lam[j] = s->lam[j + 1];
  }
  for (int i = 0; i < q; i++) s->R[i][q - 1] = 0.0;
  s->q = q - 1;
  for (int j = l; j < q - 1; j++)
  { /* zero the sub-diagonal R[j+1][j] */
    double a = s->R[j][j], b = s->R[j + 1][j];
    if (b == 0.0) continue;
    double h = hypot(a, b), c = a / h, sn = b / h;
    for (int k = j; k < q - 1; k++)
    {
      double u = s->R[j][k], v = s->R[j + 1][k];
      s->R[j][k] = c * u + sn * v;
      s->R[j + 1][k] = -sn * u + c * v;
    }
    s->R[j + 1][j] = 0.0;
    for (int i = 0; i < n; i++)
    {
      double u = s->J[i][j], v = s->J[i][j + 1];
      s->J[i][j] = c * u + sn * v;
      s->J[i][j + 1] = -sn * u + c * v;
    }
  }
}

/* directions for normal g: d = J'g, zz = |d2|^2, z = -J2 d2, r = R^-1 d1 */
static double gi_dirs(const fqo_state* s, int n, const double* g, double* d, double* z, double* r)
{
  const int q = s->q;
  for (int j = 0; j < n; j++)
  {
    double a = 0;
    for (int i = 0; i < n; i++) a += s->J[i][j] * g[i];
    d[j] = a;
  }
  double zz = 0;
  for (int j = q; j < n; j++) zz += d[j] * d[j];
  for (int i = 0; i < n; i++)
  {
    double a = 0;
    for (int j = q; j < n; j++) a += s->J[i][j] * d[j];
    z[i] = -a;
  }
  for (int k = q - 1; k >= 0; k--)
  {
    double a = d[k];
    for (int j = k + 1; j < q; j++) a -= s->R[k][j] * r[j];
    r[k] = a / s->R[k][k];
  }
  return zz;
}

/* equality g.u = h.  returns 0 ok, 1 infeasible */
static int gi_add_equality(fqo_state* s, int n, const double* g, double h)
{
  double d[FQO_NV], z[FQO_NV], r[FQO_NV];
  double sv = -h, gg = 0;
  for (int i = 0; i < n; i++) { sv += g[i] * s->x[i]; gg += g[i] * g[i]; }
  double zz = gi_dirs(s, n, g, d, z, r);
  if (zz <= FQO_EPS_DEP * gg) return fabs(sv) > FQO_TOL ? 1 : 0;
  double t = sv / zz;
  for (int i = 0; i < n; i++) s->x[i] += t * z[i];
  for (int k = 0; k < s->q; k++) s->lam[k] -= t * r[k];
  gi_add(s, n, d, -1, t);
  s->neq++;
  return 0;
}

typedef struct
{
  const fqo_model* m;
  const uint8_t* sigma;                /* assignment for segments < depth                        */
  int depth;                           /* segments with corridor rows enabled                    */
} fqo_rows;

/* most violated enabled row at u; cps evaluated with the reference's formulas from the coefficients */
static int most_violated(const fqo_rows* rs, const double* u, const uint8_t* active, double* viol_out)
{
  const fqo_model* m = rs->m;
  const int N = m->N, n = m->n, NB = n_box(m);
  double zv[FQO_NZ];
  for (int i = 0; i < 12 * N; i++)
  {
    double a = m->zc[i];
    const double* Mr = m->M[i];
    for (int j = 0; j < n; j++) a += Mr[j] * u[j];
    zv[i] = a;
  }
  double best = FQO_TOL;
  int bid = -1;
  for (int t = 0; t < N; t++)
    for (int ax = 0; ax < 3; ax++)
    {
      const double v[3] = { zv[12 * t + 6 + ax], 2 * zv[12 * t + 3 + ax], 6 * zv[12 * t + ax] };
      for (int type = 0; type < 3; type++)
        for (int sign = 0; sign < 2; sign++)
        {
          int id = ((t * 3 + ax) * 3 + type) * 2 + sign;
          if (active[id]) continue;
          double vi = (sign ? -v[type] : v[type]) - m->lim[type];
          if (vi > best) { best = vi; bid = id; }
        }
    }
  const double dt = m->dt;
  for (int t = 0; t < rs->depth; t++)
  {
    const int p = rs->sigma[t];
    double cp[4][3];
    for (int ax = 0; ax < 3; ax++)
    {
      const double a = zv[12 * t + ax], b = zv[12 * t + 3 + ax], c = zv[12 * t + 6 + ax], d = zv[12 * t + 9 + ax];
      const double bn = b * dt * dt, cn = c * dt, dn = d;
      cp[0][ax] = d;
      cp[1][ax] = (cn + 3 * dn) / 3;
      cp[2][ax] = (bn + 2 * cn + 3 * dn) / 3;
      cp[3][ax] = a * dt * dt * dt + b * dt * dt + c * dt + d;
    }
    for (int gf = m->face_ofs[p]; gf < m->face_ofs[p + 1]; gf++)
    {
      const double* ab = m->Ab + 4 * gf;
      for (int k = 0; k < 4; k++)
      {
        int id = NB + ((t * 4 + k) * m->SF + gf);
        if (active[id]) continue;
        double vi = ab[0] * cp[k][0] + ab[1] * cp[k][1] + ab[2] * cp[k][2] - ab[3];
        if (vi > best) { best = vi; bid = id; }
      }
    }
  }
  *viol_out = best;
  return bid;
}

/* run the dual method until no enabled row is violated.  1 optimal, 0 infeasible, -1 iteration cap */
static int gi_run(fqo_state* s, const fqo_rows* rs, uint8_t* active, int* iters)
{
  const fqo_model* m = rs->m;
  const int n = m->n;
  double g[FQO_NV], d[FQO_NV], z[FQO_NV], r[FQO_NV];
  for (int it = 0; it < FQO_MAX_ITER; it++)
  {
    double viol;
    int p = most_violated(rs, s->x, active, &viol);
    if (p < 0) { if (iters) *iters += it; return 1; }
    fqo_row row;
    double h = build_row(m, p, &row), c0;
    row_to_u(m, &row, g, &c0);
    h -= c0;
    double gg = 0;
    for (int i = 0; i < n; i++) gg += g[i] * g[i];
    double lam_p = 0;
    for (int inner = 0; inner < FQO_MAX_ITER; inner++)
    {
      double sv = -h;
      for (int i = 0; i < n; i++) sv += g[i] * s->x[i];
      double zz = gi_dirs(s, n, g, d, z, r);
      int dep = zz <= FQO_EPS_DEP * gg;
      double t1 = INFINITY;
      int l = -1;
      for (int k = s->neq; k < s->q; k++)
        if (r[k] > 0 && s->lam[k] / r[k] < t1) { t1 = s->lam[k] / r[k]; l = k; }
      double t2 = dep ? INFINITY : sv / zz;
      if (t1 == INFINITY && t2 == INFINITY) { if (iters) *iters += it; return 0; }
      if (t2 <= t1)
      { /* full step: the row becomes active */
        for (int i = 0; i < n; i++) s->x[i] += t2 * z[i];
        for (int k = 0; k < s->q; k++) s->lam[k] -= t2 * r[k];
        lam_p += t2;
        gi_add(s, n, d, p, lam_p);
        active[p] = 1;
        break;
      }
      if (!dep)
        for (int i = 0; i < n; i++) s->x[i] += t1 * z[i];
      for (int k = 0; k < s->q; k++) s->lam[k] -= t1 * r[k];
      lam_p += t1;
      active[s->act[l]] = 0;
      gi_drop(s, n, l);
    }
  }
  return -1;
}

/* ---------------------------------------------------------------------------------------------
 * model set-up shared by all entry points
 * ------------------------------------------------------------------------------------------- */
static int model_init(fqo_model* m, int N, int force_final, const double* x0, const double* xf, const double* lim,
                      int P, const int* face_ofs, const double* Ab, double dt)
{
  if (N < 1 || N > FQO_MAXN) return -1;
  m->N = N; m->force_final = force_final; m->P = P; m->n = 3 * N;
  m->face_ofs = face_ofs; m->Ab = Ab; m->SF = P > 0 ? face_ofs[P] : 0;
  memcpy(m->x0, x0, sizeof(m->x0)); memcpy(m->xf, xf, sizeof(m->xf)); memcpy(m->lim, lim, sizeof(m->lim));
  m->dt = dt;
  model_build_map(m);
  return 0;
}

/* final-state equalities (solverGurobi.cpp:343-356).  1 ok, 0 infeasible */
static int add_final_rows(const fqo_model* m, fqo_state* s)
{
  const int N = m->N;
  double g[FQO_NV], c0;
  for (int ax = 0; ax < 3; ax++)
  {
    fqo_row r;
    if (m->force_final)
    {
      r.nnz = 0; row_pos(&r, N - 1, m->dt, ax, 1.0); row_to_u(m, &r, g, &c0);
      if (gi_add_equality(s, m->n, g, m->xf[ax] - c0)) return 0;
    }
    r.nnz = 0; row_vel(&r, N - 1, m->dt, ax, 1.0); row_to_u(m, &r, g, &c0);
    if (gi_add_equality(s, m->n, g, m->xf[3 + ax] - c0)) return 0;
    r.nnz = 0; row_acc(&r, N - 1, m->dt, ax, 1.0); row_to_u(m, &r, g, &c0);
    if (gi_add_equality(s, m->n, g, m->xf[6 + ax] - c0)) return 0;
  }
  return 1;
}

static void write_solution(const fqo_model* m, const fqo_state* s, double* cost, double* coeffs)
{
  const int N = m->N, n = m->n;
  double c = 0;
  for (int j = 0; j < n; j++) c += s->x[j] * s->x[j];
  if (cost) *cost = c;
  if (coeffs)
    for (int i = 0; i < 12 * N; i++)
    {
      double a = m->zc[i];
      for (int j = 0; j < n; j++) a += m->M[i][j] * s->x[j];
      coeffs[i] = a;
    }
}

/* ---------------------------------------------------------------------------------------------
 * public: one fixed (dt, sigma)
 * returns 1 optimal, 0 infeasible, -1 numeric/iteration cap, -2 bad argument
 * ------------------------------------------------------------------------------------------- */
/* reusable per-thread workspace (the batch entry points solve many candidates per thread) */
typedef struct
{
  fqo_model m;
  fqo_state s;
  uint8_t* active;
  int active_cap;
} fqo_work;

static fqo_work* work_new(void)
{
  fqo_work* w = (fqo_work*)malloc(sizeof(fqo_work));
  w->active = NULL; w->active_cap = 0;
  return w;
}
static void work_free(fqo_work* w) { if (w) { free(w->active); free(w); } }

static int solve_fixed_ws(fqo_work* w, int N, int force_final, const double* x0, const double* xf, const double* lim,
                          int P, const int* face_ofs, const double* Ab, double dt, const uint8_t* sigma, double* cost,
                          double* coeffs, int* iters)
{
  fqo_model* m = &w->m;
  fqo_state* s = &w->s;
  int rc = -2;
  if (iters) *iters = 0;
  if (model_init(m, N, force_final, x0, xf, lim, P, face_ofs, Ab, dt) == 0)
  {
    gi_init(s, m->n);
    if (!add_final_rows(m, s)) rc = 0;
    else
    {
      int m_all = n_box(m) + 4 * N * m->SF;
      if (m_all < 1) m_all = 1;
      if (m_all > w->active_cap) { free(w->active); w->active = (uint8_t*)malloc(m_all); w->active_cap = m_all; }
      memset(w->active, 0, m_all);
      fqo_rows rs = { m, sigma, P > 0 ? N : 0 };
      rc = gi_run(s, &rs, w->active, iters);
      if (rc == 1) write_solution(m, s, cost, coeffs);
    }
  }
  return rc;
}

int fqo_solve_fixed(int N, int force_final, const double* x0, const double* xf, const double* lim, int P,
                    const int* face_ofs, const double* Ab, double dt, const uint8_t* sigma, double* cost,
                    double* coeffs, int* iters)
{
  fqo_work* w = work_new();
  int rc = solve_fixed_ws(w, N, force_final, x0, xf, lim, P, face_ofs, Ab, dt, sigma, cost, coeffs, iters);
  work_free(w);
  return rc;
}

/* batch of (dt, sigma) candidates for one corridor, threaded over candidates */
typedef struct
{
  int N, force_final, P;
  const double *x0, *xf, *lim;
  const int* face_ofs;
  const double* Ab;
  int n_cand;
  const double* dt;
  const uint8_t* sigma;
  uint8_t* feasible;
  double *cost, *coeffs;
  int tid, nth;
  /* heterogeneous batches: per-candidate problem index (NULL = single problem) */
  const int* prob;
  const int* poly_ofs;
} fqo_job;

static void* batch_worker(void* arg)
{
  fqo_job* j = (fqo_job*)arg;
  fqo_work* w = work_new();
  for (int i = j->tid; i < j->n_cand; i += j->nth)
  {
    double c = 0;
    int rc;
    if (j->prob)
    { /* heterogeneous batch: candidate i belongs to problem prob[i]; x0/xf/lim/poly_ofs are per problem */
      const int pr = j->prob[i];
      const int p0 = j->poly_ofs[pr], P = j->poly_ofs[pr + 1] - p0;
      int fo[64];
      const int f0 = j->face_ofs[p0];
      for (int p = 0; p <= P && p < 64; p++) fo[p] = j->face_ofs[p0 + p] - f0;
      rc = solve_fixed_ws(w, j->N, j->force_final, j->x0 + 9 * pr, j->xf + 9 * pr, j->lim + 3 * pr, P, fo,
                          j->Ab + (size_t)4 * f0, j->dt[i], j->sigma + (size_t)i * j->N, &c, NULL, NULL);
    }
    else
      rc = solve_fixed_ws(w, j->N, j->force_final, j->x0, j->xf, j->lim, j->P, j->face_ofs, j->Ab, j->dt[i],
                          j->sigma + (size_t)i * j->N, &c, j->coeffs ? j->coeffs + (size_t)i * 12 * j->N : NULL, NULL);
    j->feasible[i] = rc == 1;
    j->cost[i] = rc == 1 ? c : INFINITY;
  }
  work_free(w);
  return NULL;
}

int fqo_solve_batch(int N, int force_final, const double* x0, const double* xf, const double* lim, int P,
                    const int* face_ofs, const double* Ab, int n_cand, const double* dt, const uint8_t* sigma,
                    uint8_t* feasible, double* cost, double* coeffs, int n_threads)
{
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  fqo_job jobs[256];
  for (int t = 0; t < n_threads; t++)
  {
    fqo_job j = { N, force_final, P, x0, xf, lim, face_ofs, Ab, n_cand, dt, sigma, feasible, cost, coeffs, t,
                  n_threads, NULL, NULL };
    jobs[t] = j;
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  return 0;
}

/* many corridor problems in one call (same layout as the product's fq_solve_multi), threaded over candidates */
int fqo_solve_multi(int N, int force_final, int n_prob, const double* x0, const double* xf, const double* lim,
                    const int* poly_ofs, const int* face_ofs, const double* Ab, const int* cand_ofs, const double* dt,
                    const uint8_t* sigma, uint8_t* feasible, double* cost, int n_threads)
{
  const int n_cand = cand_ofs[n_prob];
  int* prob = (int*)malloc(sizeof(int) * (n_cand > 0 ? n_cand : 1));
  for (int j = 0; j < n_prob; j++)
    for (int i = cand_ofs[j]; i < cand_ofs[j + 1]; i++) prob[i] = j;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  fqo_job jobs[256];
  for (int t = 0; t < n_threads; t++)
  {
    fqo_job j = { N, force_final, 0, x0, xf, lim, face_ofs, Ab, n_cand, dt, sigma, feasible, cost, NULL, t,
                  n_threads, prob, poly_ofs };
    jobs[t] = j;
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  free(prob);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * public: MIQP for one dt -- minimum over ALL sigma in P^N by branch and bound
 * ------------------------------------------------------------------------------------------- */
typedef struct
{
  const fqo_model* m;
  uint8_t sigma[FQO_MAXN], best_sigma[FQO_MAXN];
  double best_cost;
  fqo_state best;
  long nodes;
  int m_all, found, numeric;
} fqo_bb;

static void bb_recurse(fqo_bb* bb, const fqo_state* parent, const uint8_t* parent_active, int depth)
{
  const fqo_model* m = bb->m;
  if (depth == m->N)
  {
    double c = 0;
    for (int j = 0; j < m->n; j++) c += parent->x[j] * parent->x[j];
    if (!bb->found || c < bb->best_cost)
    {
      bb->found = 1; bb->best_cost = c; bb->best = *parent;
      memcpy(bb->best_sigma, bb->sigma, m->N);
    }
    return;
  }
  fqo_state* s = (fqo_state*)malloc(sizeof(fqo_state));
  uint8_t* active = (uint8_t*)malloc(bb->m_all);
  for (int p = 0; p < m->P; p++)
  {
    *s = *parent;
    memcpy(active, parent_active, bb->m_all);
    bb->sigma[depth] = (uint8_t)p;
    fqo_rows rs = { m, bb->sigma, depth + 1 };
    bb->nodes++;
    int rc = gi_run(s, &rs, active, NULL);
    if (rc < 0) bb->numeric++;
    if (rc != 1) continue;
    double c = 0;
    for (int j = 0; j < m->n; j++) c += s->x[j] * s->x[j];
    if (bb->found && c >= bb->best_cost) continue;   /* relaxation bound: children only add rows */
    bb_recurse(bb, s, active, depth + 1);
  }
  free(active); free(s);
}

int fqo_solve_miqp(int N, int force_final, const double* x0, const double* xf, const double* lim, int P,
                   const int* face_ofs, const double* Ab, double dt, uint8_t* sigma_out, double* cost,
                   double* coeffs, long* nodes)
{
  fqo_model* m = (fqo_model*)malloc(sizeof(fqo_model));
  fqo_state* s = (fqo_state*)malloc(sizeof(fqo_state));
  fqo_bb* bb = (fqo_bb*)calloc(1, sizeof(fqo_bb));
  int rc = -2;
  uint8_t* active = NULL;
  if (nodes) *nodes = 0;
  if (model_init(m, N, force_final, x0, xf, lim, P, face_ofs, Ab, dt) == 0)
  {
    gi_init(s, m->n);
    if (!add_final_rows(m, s)) rc = 0;
    else
    {
      bb->m = m; bb->m_all = n_box(m) + 4 * N * m->SF; if (bb->m_all < 1) bb->m_all = 1;
      active = (uint8_t*)calloc(bb->m_all, 1);
      fqo_rows rs = { m, bb->sigma, 0 };
      rc = gi_run(s, &rs, active, NULL);       /* root: boxes only (also the whole answer when P == 0) */
      if (rc == 1)
      {
        if (P == 0) { bb->found = 1; bb->best = *s; }
        else bb_recurse(bb, s, active, 0);
        rc = bb->found ? 1 : (bb->numeric ? -1 : 0);
        if (bb->found)
        {
          write_solution(m, &bb->best, cost, coeffs);
          if (sigma_out) memcpy(sigma_out, bb->best_sigma, N);
        }
      }
      if (nodes) *nodes = bb->nodes;
    }
  }
  free(active); free(bb); free(s); free(m);
  return rc;
}

/* ---------------------------------------------------------------------------------------------
 * getDTInitial (solverGurobi.cpp:659-759) incl. the float temporaries; MinPositiveElement
 * (solverGurobi_utils.hpp:19-32): smallest root > 0, or 0 when there is none.
 * The reference finds roots with Eigen's companion-matrix PolynomialSolver; here they come from the closed
 * forms polished by Newton (they differ by a few ulp of double and are then rounded to float).
 * ------------------------------------------------------------------------------------------- */
static int real_roots_quadratic(double c0, double c1, double c2, double* r)
{ /* c2 t^2 + c1 t + c0 */
  double disc = c1 * c1 - 4 * c2 * c0;
  if (disc < 0) return 0;
  double sq = sqrt(disc);
  double qv = -0.5 * (c1 + (c1 >= 0 ? sq : -sq));
  int k = 0;
  r[k++] = qv / c2;
  if (qv != 0) r[k++] = c0 / qv; else r[k++] = 0.0;
  return k;
}

static int real_roots_cubic(double c0, double c1, double c2, double c3, double* r)
{ /* c3 t^3 + c2 t^2 + c1 t + c0 */
  double a = c2 / c3, b = c1 / c3, c = c0 / c3;
  double Q = (a * a - 3 * b) / 9, Rr = (2 * a * a * a - 9 * a * b + 27 * c) / 54;
  int k = 0;
  if (Rr * Rr < Q * Q * Q)
  {
    double th = acos(Rr / sqrt(Q * Q * Q)), sq = -2 * sqrt(Q);
    r[k++] = sq * cos(th / 3) - a / 3;
    r[k++] = sq * cos((th + 2 * M_PI) / 3) - a / 3;
    r[k++] = sq * cos((th - 2 * M_PI) / 3) - a / 3;
  }
  else
  {
    double A = -copysign(cbrt(fabs(Rr) + sqrt(Rr * Rr - Q * Q * Q)), Rr);
    double B = A != 0 ? Q / A : 0;
    r[k++] = A + B - a / 3;
    if (Rr * Rr == Q * Q * Q && Q != 0) { r[k++] = -0.5 * (A + B) - a / 3; }
  }
  for (int i = 0; i < k; i++)
    for (int it = 0; it < 3; it++)
    {
      double t = r[i], f = ((c3 * t + c2) * t + c1) * t + c0, fp = (3 * c3 * t + 2 * c2) * t + c1;
      if (fp != 0 && isfinite(f / fp)) r[i] = t - f / fp;
    }
  return k;
}

static double min_positive(const double* v, int n)
{
  double best = 0;
  int found = 0;
  for (int i = 0; i < n; i++)
    if (v[i] > 0 && (!found || v[i] < best)) { best = v[i]; found = 1; }
  return best;
}

double fqo_dt_initial(const double* x0, const double* xf, const double* lim, int N)
{
  const double v_max = lim[0], a_max = lim[1], j_max = lim[2];
  float tv[3], ta[3], tj[3];
  for (int i = 0; i < 3; i++)
  {
    tv[i] = (float)(fabs(xf[i] - x0[i]) / v_max);
    float jerk = (float)(copysign(1, xf[i] - x0[i]) * j_max);
    float accel = (float)(copysign(1, xf[i] - x0[i]) * a_max);
    float a0 = (float)x0[6 + i], v0 = (float)x0[3 + i];
    double r[3];
    int k = real_roots_cubic(x0[i] - xf[i], v0, a0 / 2.0, jerk / 6.0, r);
    tj[i] = (float)min_positive(r, k);
    k = real_roots_quadratic(x0[i] - xf[i], v0, 0.5 * accel, r);
    ta[i] = (float)min_positive(r, k);
  }
  float mx = tv[0];
  for (int i = 0; i < 3; i++)
  {
    if (tv[i] > mx) mx = tv[i];
    if (ta[i] > mx) mx = ta[i];
    if (tj[i] > mx) mx = tj[i];
  }
  double dt_initial = (double)(float)(mx / (float)N);   /* float / int -> float arithmetic, as in the reference (:751) */
  if (dt_initial > 10000) dt_initial = 0;
  return dt_initial;
}

/* ---------------------------------------------------------------------------------------------
 * genNewTraj sweep (solverGurobi.cpp:426-477).  sigma_list == NULL -> branch and bound over all sigma.
 * outputs mirror the reference's public members: dt_, factor_that_worked_, trials_.
 * ------------------------------------------------------------------------------------------- */
int fqo_gen_new_traj(int N, int force_final, const double* x0, const double* xf, const double* lim, int P,
                     const int* face_ofs, const double* Ab, double DC, double factor_initial, double factor_final,
                     double factor_increment, int n_sigma, const uint8_t* sigma_list, double* dt_out,
                     double* factor_out, int* trials_out, uint8_t* sigma_out, double* cost_out, double* coeffs)
{
  int solved = 0, trials = 0;
  double dt = 0;
  double dti = fqo_dt_initial(x0, xf, lim, N);
  for (double f = factor_initial; f <= factor_final && !solved; f = f + factor_increment)
  {
    trials++;
    dt = f * fmax(dti, 2 * DC);        /* findDT (:494-497) */
    if (P == 0 || sigma_list == NULL)
    {
      int rc = fqo_solve_miqp(N, force_final, x0, xf, lim, P, face_ofs, Ab, dt, sigma_out, cost_out, coeffs, NULL);
      solved = rc == 1;
    }
    else
    {
      double best = INFINITY;
      double* tmp = (double*)malloc(sizeof(double) * 12 * N);
      for (int k = 0; k < n_sigma; k++)
      {
        double c;
        int rc = fqo_solve_fixed(N, force_final, x0, xf, lim, P, face_ofs, Ab, dt, sigma_list + (size_t)k * N, &c,
                                 tmp, NULL);
        if (rc == 1 && c < best)
        {
          best = c; solved = 1;
          if (cost_out) *cost_out = c;
          if (coeffs) memcpy(coeffs, tmp, sizeof(double) * 12 * N);
          if (sigma_out) memcpy(sigma_out, sigma_list + (size_t)k * N, N);
        }
      }
      free(tmp);
    }
    if (solved && factor_out) *factor_out = f;
  }
  if (dt_out) *dt_out = dt;
  if (trials_out) *trials_out = trials;
  return solved;
}

/* resetX (solverGurobi.cpp:382-388): number of samples */
int fqo_num_samples(int N, double dt, double DC)
{
  int size = (int)(N)*dt / DC;
  return size < 2 ? 2 : size;
}

/* fillX (solverGurobi.cpp:122-168): out[i][0..11] = pos(3) vel(3) accel(3) jerk(3) */
void fqo_fill_x(int N, const double* coeffs, double dt, double DC, int n_samples, double* out)
{
  double t = 0;
  int interval = 0;
  for (int i = 0; i < n_samples; i++)
  {
    t = t + DC;
    if (t > dt * (interval + 1)) interval = (interval + 1 < N - 1) ? interval + 1 : N - 1;
    double tau = t - interval * dt;
    const double* x = coeffs + 12 * interval;
    for (int ax = 0; ax < 3; ax++)
    {
      out[12 * i + ax] = x[ax] * tau * tau * tau + x[3 + ax] * tau * tau + x[6 + ax] * tau + x[9 + ax];
      out[12 * i + 3 + ax] = 3 * x[ax] * tau * tau + 2 * x[3 + ax] * tau + x[6 + ax];
      out[12 * i + 6 + ax] = 6 * x[ax] * tau + 2 * x[3 + ax];
      out[12 * i + 9 + ax] = 6 * x[ax];
    }
  }
  for (int k = 3; k < 12; k++) out[12 * (n_samples - 1) + k] = 0.0;   /* :165-167 */
}

int fqo_abi_version(void) { return 1; }

/* row tolerance of every later solve (default 1e-8; Gurobi's FeasibilityTol default is 1e-6, solverGurobi.cpp:479-487
 * sets no tolerance parameter) */
void fqo_set_row_tol(double tol) { if (tol >= 0) g_row_tol = tol; }
double fqo_get_row_tol(void) { return g_row_tol; }

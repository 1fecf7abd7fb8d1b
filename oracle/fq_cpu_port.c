/* TEST / BASELINE INFRASTRUCTURE ONLY -- a TUNED CPU port of the GPU kernel's algorithm, the honest CPU arm of bench.py.
 *
 * oracle/fq_oracle.c is the literal restatement (rows written over the 12N coefficients, dense full-J Goldfarb-Idnani,
 * entering row = largest violation): it is the checker.  This file is what a CPU implementation of the SAME method as
 * faster_b200/csrc/fq_kernels_t.cuh looks like when it is written for speed:
 *   - normalised time and eliminated final-state equalities (plan tables per (N, force_final), built once): every
 *     model row is a constant functional of the unknowns w, Y = Yeq + TZ w; min |w|^2 subject to inequality rows only
 *     (model rows: reference solverGurobi.cpp:113-119 cost, :332-380,:499-524 equalities, :390-407 boxes, :249-287 and
 *     :833-862 corridor rows of the four Bezier control points);
 *   - dual active set with the thin factorisation J1 R of the active normals, entering row by normalised violation
 *     (distance to the row's hyperplane), exactly the GPU kernel's rules;
 *   - flat arrays, no allocation per candidate, loops the compiler vectorises (-O3 -march=native), a persistent
 *     pthread pool that claims candidates dynamically (chunks of 8 from an atomic counter).
 * Results are checked against fq_oracle.c in tests/test_cpu_port.py (flags equal, costs to 1e-9).
 * Nothing in the product (faster_b200/, include/) links or loads this.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 16
#define MAXNZ (MAXN - 2)
#define MAXNW (3 * MAXNZ)
#define MAXNY (6 * MAXN + 1)
#define MAXROWS 4096
#define EPS_DEP 1e-18
#define ZZ_FLOOR 1e-30
#define MAX_ITERS 400

typedef long double ld;

typedef struct
{
  int N, ff, ne, nz, nw, NY, ok;
  double TZ[MAXNY][MAXNZ];      /* row y: functional over w of one axis */
  double T0[MAXNY][6];          /* (P0,V0,A0) and the eliminated right-hand sides */
  double FT[3][3];
  double SY[MAXNY];             /* 1 / |TZ[y]| */
} plan_t;

static plan_t g_plans[2 * (MAXN + 1)];
static pthread_mutex_t g_plan_mu = PTHREAD_MUTEX_INITIALIZER;
static double g_tol = 1e-8;

void fqc_set_row_tol(double t) { if (t >= 0) g_tol = t; }

/* ---- plan tables: triple integrator in normalised time (ubar = u dt^3, V = v dt, A = a dt^2) ---------------------- */
static ld cP(int k) { return (3.0L * k * k + 3.0L * k + 1.0L) / 6.0L; }
static ld cV(int k) { return k + 0.5L; }

static void functional(int N, int y, ld* c, ld f[3])
{ /* row y of Y as (coefficients over ubar_0..N-1, response to (P0,V0,A0)) */
  for (int i = 0; i < N; i++) c[i] = 0;
  f[0] = f[1] = f[2] = 0;
#define ADDP(t, w) do { for (int s = 0; s < (t); s++) c[s] += (w) * cP((t) - 1 - s); f[0] += (w); f[1] += (w) * (t); f[2] += (w) * (ld)(t) * (t) / 2.0L; } while (0)
#define ADDV(t, w) do { for (int s = 0; s < (t); s++) c[s] += (w) * cV((t) - 1 - s); f[1] += (w); f[2] += (w) * (t); } while (0)
#define ADDA(t, w) do { for (int s = 0; s < (t); s++) c[s] += (w); f[2] += (w); } while (0)
  if (y <= N) ADDP(y, 1.0L);
  else if (y < 2 * N + 1) ADDV(y - (N + 1), 1.0L);
  else if (y < 3 * N + 1) ADDA(y - (2 * N + 1), 1.0L);
  else if (y < 4 * N + 1) c[y - (3 * N + 1)] = 1;
  else if (y < 5 * N + 1) { int t = y - (4 * N + 1); ADDP(t, 1.0L); ADDV(t, 1.0L / 3.0L); }
  else { int t = y - (5 * N + 1); ADDP(t, 1.0L); ADDV(t, 2.0L / 3.0L); ADDA(t, 1.0L / 6.0L); }
}

static const plan_t* get_plan(int N, int ff)
{
  const int ne = ff ? 3 : 2;
  if (N < ne || N > MAXN) return NULL;
  plan_t* p = &g_plans[2 * N + (ff ? 1 : 0)];
  if (p->ok) return p;
  pthread_mutex_lock(&g_plan_mu);
  if (!p->ok)
  {
    const int nz = N - ne, NY = 6 * N + 1;
    ld C[3][MAXN], Q[MAXN][MAXN], A[MAXN][3], Ep[MAXN][3];
    memset(p, 0, sizeof(*p));
    p->N = N; p->ff = ff; p->ne = ne; p->nz = nz; p->nw = 3 * nz; p->NY = NY;
    int e = 0;
    if (ff) { for (int s = 0; s < N; s++) C[e][s] = cP(N - 1 - s); p->FT[e][0] = 1; p->FT[e][1] = N; p->FT[e][2] = (double)N * N / 2.0; e++; }
    for (int s = 0; s < N; s++) C[e][s] = cV(N - 1 - s);
    p->FT[e][1] = 1; p->FT[e][2] = N; e++;
    for (int s = 0; s < N; s++) C[e][s] = 1;
    p->FT[e][2] = 1;
    /* Householder QR of C' (N x ne): columns ne..N-1 of Q span the null space of the terminal rows */
    for (int i = 0; i < N; i++) { for (int k = 0; k < ne; k++) A[i][k] = C[k][i]; for (int j = 0; j < N; j++) Q[i][j] = i == j; }
    for (int k = 0; k < ne; k++)
    {
      ld nrm = 0, v[MAXN], vv = 0;
      for (int i = k; i < N; i++) nrm += A[i][k] * A[i][k];
      nrm = sqrtl(nrm);
      for (int i = 0; i < N; i++) v[i] = i >= k ? A[i][k] : 0;
      v[k] -= A[k][k] >= 0 ? -nrm : nrm;
      for (int i = k; i < N; i++) vv += v[i] * v[i];
      if (vv == 0) continue;
      for (int c = 0; c < ne; c++) { ld s = 0; for (int i = k; i < N; i++) s += v[i] * A[i][c]; s = 2 * s / vv; for (int i = k; i < N; i++) A[i][c] -= s * v[i]; }
      for (int r = 0; r < N; r++) { ld s = 0; for (int i = k; i < N; i++) s += Q[r][i] * v[i]; s = 2 * s / vv; for (int i = k; i < N; i++) Q[r][i] -= s * v[i]; }
    }
    for (int i = 0; i < N; i++)       /* minimum-norm particular solution: Ep = Q1 R^-T */
      for (int b = ne - 1; b >= 0; b--)
      {
        ld s = Q[i][b];
        for (int a = b + 1; a < ne; a++) s -= Ep[i][a] * A[b][a];
        Ep[i][b] = s / A[b][b];
      }
    for (int y = 0; y < NY; y++)
    {
      ld c[MAXN], f[3], nrm = 0, cn = 0, tz[MAXNZ];
      functional(N, y, c, f);
      for (int m = 0; m < nz; m++) { ld s = 0; for (int i = 0; i < N; i++) s += c[i] * Q[i][ne + m]; tz[m] = s; nrm += s * s; }
      for (int i = 0; i < N; i++) cn += c[i] * c[i];
      const int zero = nrm <= 1e-24L * (cn > 1 ? cn : 1);
      double s2 = 0;
      for (int m = 0; m < nz; m++) { p->TZ[y][m] = zero ? 0.0 : (double)tz[m]; s2 += p->TZ[y][m] * p->TZ[y][m]; }
      p->SY[y] = s2 > 1e-30 ? 1.0 / sqrt(s2) : 1e15;
      for (int k = 0; k < 3; k++) p->T0[y][k] = (double)f[k];
      for (int k = 0; k < ne; k++) { ld s = 0; for (int i = 0; i < N; i++) s += c[i] * Ep[i][k]; p->T0[y][3 + k] = (double)s; }
    }
    p->ok = 1;
  }
  pthread_mutex_unlock(&g_plan_mu);
  return p;
}

/* ---- one candidate ------------------------------------------------------------------------------------------------ */
typedef struct
{
  double Yeq[3][MAXNY], Y[3][MAXNY];
  double J[MAXNW][MAXNW];       /* column k = k-th basis vector, stored as J[k][0..nw) (contiguous per column) */
  double R[MAXNW][MAXNW];       /* R[k][j], j <= k: column k of the upper triangular factor */
  double w[MAXNW], lam[MAXNW], g[MAXNW], d[MAXNW], z[MAXNW], r[MAXNW];
  int aseg[MAXNW];              /* segment (1-based; 0 = box row) each active row came from */
  int rseg[MAXROWS];            /* segment of each corridor row */
  /* corridor rows of the candidate: normal, offset, the up-to-four Y rows they apply to */
  double ra[MAXROWS][4];
  int ry[MAXROWS][4], rn[MAXROWS];
  int n_rows;
} work_t;

static void update_Y(const plan_t* p, work_t* k)
{
  const int NY = p->NY, nz = p->nz;
  for (int ax = 0; ax < 3; ax++)
  {
    const double* wa = k->w + ax * nz;
    for (int y = 0; y < NY; y++)
    {
      double acc = k->Yeq[ax][y];
      for (int m = 0; m < nz; m++) acc += p->TZ[y][m] * wa[m];
      k->Y[ax][y] = acc;
    }
  }
}

/* ---- infeasibility certificates shared between the candidates of one problem -------------------------------------------
 * When the iteration stops with "infeasible", the entering row e and the active rows k with r_k < 0 are a Farkas
 * certificate: g_e = sum r_k n_k with r <= 0, every active row is tight and row e is violated, so NO point satisfies
 * those rows together -- whatever the other rows are.  Every other candidate of the same problem with the same dt and
 * the same polytope on the segments those rows belong to (box rows belong to none) contains the same rows and is
 * infeasible too.  A time-allocation sweep starts at an optimistic dt, so whole runs of candidates die by the boxes
 * alone, or by the first segments' rows: they are answered from the memo instead of being re-proved.  Certificates are
 * only recorded when the violation exceeds MEMO_MARGIN (1 + sum |r_k|), far from the tolerance band. */
#ifndef MEMO_NB
#define MEMO_NB 16            /* buckets per problem, by a hash of dt */
#endif
#ifndef MEMO_BE
#define MEMO_BE 8             /* certificates per bucket */
#endif
#define MEMO_MARGIN 1e-5
typedef struct { uint64_t dt_bits, sigpack; uint32_t mask; } memo_entry;
typedef struct { atomic_int n[MEMO_NB], ready[MEMO_NB][MEMO_BE]; memo_entry e[MEMO_NB][MEMO_BE]; } memo_t;
static inline int memo_bucket(uint64_t dt_bits) { return (int)((dt_bits * 0x9E3779B97F4A7C15ull) >> 40) % MEMO_NB; }
static int g_memo_on = 1;
static int g_memo_lookup = 1;      /* analysis: 0 = record certificates but never answer from the memo */
void fqc_set_memo_lookup(int on) { g_memo_lookup = on; }
static int32_t* g_cert_out = NULL;   /* analysis only: per candidate, the certificate's segment mask (-1: none) */
void fqc_set_cert_out(int32_t* p) { g_cert_out = p; }
static __thread long g_cur_cand = -1;
void fqc_set_memo(int on) { g_memo_on = on; }
static atomic_long g_memo_hits, g_memo_inserts;
void fqc_memo_stats(long* hits, long* inserts, int reset)
{
  if (hits) *hits = atomic_load(&g_memo_hits);
  if (inserts) *inserts = atomic_load(&g_memo_inserts);
  if (reset) { atomic_store(&g_memo_hits, 0); atomic_store(&g_memo_inserts, 0); }
}
static inline uint64_t nibble_mask(uint32_t m)
{
  uint64_t r = 0;
  for (int t = 0; t < 16; t++) if (m >> t & 1) r |= 0xfull << (4 * t);
  return r;
}

/* returns 1 optimal, 0 infeasible, -1 iteration cap / NaN */
static int solve_one(const plan_t* p, work_t* k, const double* x0, const double* xf, const double* lim, double dt, int P,
                     const int* face_ofs, const double* Ab, const uint8_t* sigma, double* cost, double* coeffs, int* iters,
                     memo_t* memo)
{
  const int N = p->N, ne = p->ne, nz = p->nz, nw = p->nw, NY = p->NY;
  const double tol = g_tol, dt2 = dt * dt, dt3 = dt2 * dt;
  const double inv[3] = { 1.0 / dt, 1.0 / dt2, 1.0 / dt3 };
  if (!(dt > 0 && dt < 1e100)) { if (iters) *iters = -1; return -1; }
  uint64_t dt_bits = 0, sigpack = 0;
  memcpy(&dt_bits, &dt, 8);
  if (memo && P > 0 && P <= 16)
  {
    for (int t = 0; t < N; t++) sigpack |= (uint64_t)(sigma[t] < P ? sigma[t] : P - 1) << (4 * t);
    const int bk = memo_bucket(dt_bits);
    const int n = atomic_load_explicit(&memo->n[bk], memory_order_acquire);
    for (int i = 0; g_memo_lookup && i < n && i < MEMO_BE; i++)
      if (atomic_load_explicit(&memo->ready[bk][i], memory_order_acquire) && memo->e[bk][i].dt_bits == dt_bits &&
          ((memo->e[bk][i].sigpack ^ sigpack) & nibble_mask(memo->e[bk][i].mask)) == 0)
      {
        atomic_fetch_add(&g_memo_hits, 1);
        if (iters) *iters = 0;
        if (cost) *cost = INFINITY;
        if (coeffs) memset(coeffs, 0, sizeof(double) * 12 * (size_t)N);
        return 0;
      }
  }
  else memo = NULL;
  for (int ax = 0; ax < 3; ax++)
  {
    const double s0 = x0[ax], s1 = x0[3 + ax] * dt, s2 = x0[6 + ax] * dt2;
    double rhs[3], tgt[3];
    int e = 0;
    if (p->ff) tgt[e++] = xf[ax];
    tgt[e++] = xf[3 + ax] * dt; tgt[e++] = xf[6 + ax] * dt2;
    for (int q = 0; q < ne; q++) rhs[q] = tgt[q] - (p->FT[q][0] * s0 + p->FT[q][1] * s1 + p->FT[q][2] * s2);
    for (int y = 0; y < NY; y++)
    {
      double v = p->T0[y][0] * s0 + p->T0[y][1] * s1 + p->T0[y][2] * s2;
      for (int q = 0; q < ne; q++) v += p->T0[y][3 + q] * rhs[q];
      k->Yeq[ax][y] = v; k->Y[ax][y] = v;
    }
  }
  /* corridor rows: control points 1, 2, 3 of segment t always; control point 0 only where the polytope changes (it is
     the previous segment's control point 3 otherwise) */
  int nr = 0;
  if (P > 0)
    for (int t = 0; t < N; t++)
    {
      int pp = sigma[t] < P ? sigma[t] : P - 1;
      const int need0 = t == 0 || sigma[t - 1] != sigma[t];
      for (int f = face_ofs[pp]; f < face_ofs[pp + 1]; f++)
      {
        if (nr >= MAXROWS) { if (iters) *iters = -2; return -1; }
        k->ra[nr][0] = Ab[4 * f]; k->ra[nr][1] = Ab[4 * f + 1]; k->ra[nr][2] = Ab[4 * f + 2]; k->ra[nr][3] = Ab[4 * f + 3];
        k->ry[nr][0] = 4 * N + 1 + t; k->ry[nr][1] = 5 * N + 1 + t; k->ry[nr][2] = t + 1; k->ry[nr][3] = t;
        k->rn[nr] = need0 ? 4 : 3;
        k->rseg[nr] = t + 1;
        nr++;
      }
    }
  k->n_rows = nr;
  for (int i = 0; i < nw; i++) k->w[i] = 0;
  int q = 0, it = 0, status = -2;
  const double bthr[3] = { (lim[0] + tol) * dt, (lim[1] + tol) * dt2, (lim[2] + tol) * dt3 };
  while (status == -2)
  {
    /* ---- entering row: farthest beyond its hyperplane in w-space */
    double best = 0, wv[3] = { 0, 0, 0 }, h = 0;
    int by = -1, bseg = 0;
    for (int typ = 0; typ < 3; typ++)
      for (int ax = 0; ax < 3; ax++)
      {
        const double* Ya = k->Y[ax] + (typ + 1) * N + 1;
        const double* S = p->SY + (typ + 1) * N + 1;
        for (int t = 0; t < N; t++)
        {
          const double rk = (fabs(Ya[t]) - bthr[typ]) * S[t];
          if (rk > best)
          {
            best = rk; by = (typ + 1) * N + 1 + t; h = lim[typ]; bseg = 0;
            wv[0] = wv[1] = wv[2] = 0; wv[ax] = Ya[t] > 0 ? inv[typ] : -inv[typ];
          }
        }
      }
    for (int i = 0; i < nr; i++)
    {
      const double a0 = k->ra[i][0], a1 = k->ra[i][1], a2 = k->ra[i][2], bb = k->ra[i][3] + tol;
      for (int c = 0; c < k->rn[i]; c++)
      {
        const int y = k->ry[i][c];
        const double rk = (a0 * k->Y[0][y] + a1 * k->Y[1][y] + a2 * k->Y[2][y] - bb) * p->SY[y];
        if (rk > best) { best = rk; by = y; wv[0] = a0; wv[1] = a1; wv[2] = a2; h = k->ra[i][3]; bseg = k->rseg[i]; }
      }
    }
    if (by < 0) { status = 1; break; }
    if (best != best) { status = -1; break; }
    double gg = 0;
    for (int ax = 0; ax < 3; ax++)
      for (int m = 0; m < nz; m++) { const double gv = wv[ax] * p->TZ[by][m]; k->g[ax * nz + m] = gv; gg += gv * gv; }
    double lam_p = 0;
    for (;;)
    {
      if (++it > MAX_ITERS) { status = -1; break; }
      const double viol = wv[0] * k->Y[0][by] + wv[1] * k->Y[1][by] + wv[2] * k->Y[2][by] - h;
      /* d1 = J1' g, z = -(g - J1 d1) */
      for (int i = 0; i < nw; i++) k->z[i] = -k->g[i];
      for (int c = 0; c < q; c++)
      {
        double s = 0;
        const double* Jc = k->J[c];
        for (int i = 0; i < nw; i++) s += Jc[i] * k->g[i];
        k->d[c] = s;
        for (int i = 0; i < nw; i++) k->z[i] += Jc[i] * s;
      }
      double zz = 0;
      for (int i = 0; i < nw; i++) zz += k->z[i] * k->z[i];
      if (q > 0 && zz < 0.01 * gg && zz > fmax(EPS_DEP * gg, ZZ_FLOOR))
      { /* re-orthogonalise once */
        for (int c = 0; c < q; c++)
        {
          double s = 0;
          const double* Jc = k->J[c];
          for (int i = 0; i < nw; i++) s += Jc[i] * k->z[i];
          k->d[c] -= s;
          for (int i = 0; i < nw; i++) k->z[i] -= Jc[i] * s;
        }
        zz = 0;
        for (int i = 0; i < nw; i++) zz += k->z[i] * k->z[i];
      }
      /* r = R^-1 d1 */
      for (int c = q - 1; c >= 0; c--)
      {
        double s = k->d[c];
        for (int j = c + 1; j < q; j++) s -= k->R[j][c] * k->r[j];
        k->r[c] = s / k->R[c][c];
      }
      const int dep = zz <= fmax(EPS_DEP * gg, ZZ_FLOOR) || q >= nw;
      double t1 = INFINITY;
      int l = -1;
      for (int c = 0; c < q; c++)
        if (k->r[c] > 0) { const double ratio = k->lam[c] / k->r[c]; if (ratio < t1) { t1 = ratio; l = c; } }
      const double t2 = dep ? INFINITY : viol / zz;
      if (t1 == INFINITY && t2 == INFINITY)
      {
        status = 0;
        if (memo)
        { /* record the certificate: rows with a negative multiplier + the entering row */
          double sr = 0;
          uint32_t mask = bseg ? 1u << (bseg - 1) : 0u;
          for (int c = 0; c < q; c++)
            if (k->r[c] < 0) { sr -= k->r[c]; if (k->aseg[c]) mask |= 1u << (k->aseg[c] - 1); }
          if (g_cert_out && g_cur_cand >= 0) g_cert_out[g_cur_cand] = viol > MEMO_MARGIN * (1.0 + sr) ? (int32_t)mask : -1;
          if (viol > MEMO_MARGIN * (1.0 + sr))
          {
            const int bk = memo_bucket(dt_bits);
            const int slot = atomic_fetch_add(&memo->n[bk], 1);
            if (slot < MEMO_BE)
            {
              memo->e[bk][slot].dt_bits = dt_bits; memo->e[bk][slot].sigpack = sigpack; memo->e[bk][slot].mask = mask;
              atomic_store_explicit(&memo->ready[bk][slot], 1, memory_order_release);
              atomic_fetch_add(&g_memo_inserts, 1);
            }
            else atomic_store(&memo->n[bk], MEMO_BE);
          }
        }
        break;
      }
      if (t2 <= t1)
      { /* full step: the row becomes active */
        const double nrm = sqrt(zz), rn = 1.0 / nrm;
        for (int i = 0; i < nw; i++) { k->w[i] += t2 * k->z[i]; k->J[q][i] = -k->z[i] * rn; }
        for (int c = 0; c < q; c++) { k->lam[c] -= t2 * k->r[c]; k->R[q][c] = k->d[c]; }
        k->R[q][q] = nrm; k->lam[q] = lam_p + t2; k->aseg[q] = bseg;
        q++;
        update_Y(p, k);
        break;
      }
      if (l < 0) { status = -1; break; }
      if (!dep) for (int i = 0; i < nw; i++) k->w[i] += t1 * k->z[i];
      for (int c = 0; c < q; c++) k->lam[c] -= t1 * k->r[c];
      lam_p += t1;
      /* drop active element l: shift columns, restore the triangle with Givens rotations (applied to J's columns too) */
      for (int c = l; c < q - 1; c++)
      {
        memcpy(k->R[c], k->R[c + 1], sizeof(double) * (size_t)(c + 2));
        k->lam[c] = k->lam[c + 1];
        k->aseg[c] = k->aseg[c + 1];
      }
      for (int j = l; j < q - 1; j++)
      {
        const double pv = k->R[j][j], sv = k->R[j][j + 1];
        const double hh = sqrt(pv * pv + sv * sv);
        double cs = 1, sn = 0;
        if (hh > 0) { cs = pv / hh; sn = sv / hh; }
        for (int c = j; c < q - 1; c++)
        {
          const double u = k->R[c][j], v = k->R[c][j + 1];
          k->R[c][j] = cs * u + sn * v; k->R[c][j + 1] = cs * v - sn * u;
        }
        double *Ja = k->J[j], *Jb = k->J[j + 1];
        for (int i = 0; i < nw; i++) { const double u = Ja[i], v = Jb[i]; Ja[i] = cs * u + sn * v; Jb[i] = cs * v - sn * u; }
      }
      q--;
      if (!dep) update_Y(p, k);
    }
  }
  if (iters) *iters = status == -1 ? -it : it;
  if (status != 1) { if (cost) *cost = INFINITY; if (coeffs) memset(coeffs, 0, sizeof(double) * 12 * (size_t)N); return status; }
  double cp = 0;
  for (int ax = 0; ax < 3; ax++)
    for (int t = 0; t < N; t++) { const double u = k->Y[ax][3 * N + 1 + t]; cp += u * u; }
  if (cost) *cost = cp * inv[2] * inv[2];
  if (coeffs)
    for (int t = 0; t < N; t++)
      for (int ax = 0; ax < 3; ax++)
      {
        coeffs[12 * t + ax] = k->Y[ax][3 * N + 1 + t] * inv[2] / 6.0;
        coeffs[12 * t + 3 + ax] = k->Y[ax][2 * N + 1 + t] * inv[1] / 2.0;
        coeffs[12 * t + 6 + ax] = k->Y[ax][N + 1 + t] * inv[0];
        coeffs[12 * t + 9 + ax] = k->Y[ax][t];
      }
  return 1;
}

/* ---- persistent pool with dynamic claiming -------------------------------------------------------------------------- */
typedef struct
{
  const plan_t* plan;
  int n_prob;
  const double *x0, *xf, *lim, *Ab, *dt;
  const int *poly_ofs, *face_ofs, *cand_ofs;
  const uint8_t* sigma;
  uint8_t* feasible;
  double *cost, *coeffs;
  int32_t* iters;
  long n_cand;
  memo_t* memo;                 /* n_prob entries or NULL */
} job_t;

static struct
{
  pthread_t th[512];
  int n_threads, started;
  pthread_mutex_t mu;
  pthread_cond_t go, done;
  long generation;
  int active, quit;
  job_t job;
  atomic_long next;
} g_pool = { .mu = PTHREAD_MUTEX_INITIALIZER, .go = PTHREAD_COND_INITIALIZER, .done = PTHREAD_COND_INITIALIZER };

static void run_job(const job_t* jb, work_t* wk)
{
  const int N = jb->plan->N;
  int prob = 0;
  for (;;)
  {
    const long c0 = atomic_fetch_add(&g_pool.next, 8);
    if (c0 >= jb->n_cand) break;
    const long c1 = c0 + 8 < jb->n_cand ? c0 + 8 : jb->n_cand;
    for (long c = c0; c < c1; c++)
    {
      while (prob + 1 < jb->n_prob && jb->cand_ofs[prob + 1] <= c) prob++;
      while (prob > 0 && jb->cand_ofs[prob] > c) prob--;
      const int p0 = jb->poly_ofs[prob], P = jb->poly_ofs[prob + 1] - p0;
      int fo[64];
      const int f0 = jb->face_ofs[p0];
      for (int i = 0; i <= P && i < 64; i++) fo[i] = jb->face_ofs[p0 + i] - f0;
      double cst = INFINITY;
      int it = 0;
      g_cur_cand = c;
      const int st = solve_one(jb->plan, wk, jb->x0 + 9 * prob, jb->xf + 9 * prob, jb->lim + 3 * prob, jb->dt[c], P, fo, jb->Ab + 4 * (size_t)f0,
                               jb->sigma ? jb->sigma + (size_t)c * N : NULL, &cst, jb->coeffs ? jb->coeffs + (size_t)c * 12 * N : NULL, &it,
                               jb->memo && jb->sigma ? jb->memo + prob : NULL);
      jb->feasible[c] = st == 1;
      jb->cost[c] = st == 1 ? cst : INFINITY;
      if (jb->iters) jb->iters[c] = it;
    }
  }
}

static void* worker(void* arg)
{
  (void)arg;
  work_t* wk = (work_t*)malloc(sizeof(work_t));
  long seen = 0;
  for (;;)
  {
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.generation == seen && !g_pool.quit) pthread_cond_wait(&g_pool.go, &g_pool.mu);
    if (g_pool.quit) { pthread_mutex_unlock(&g_pool.mu); break; }
    seen = g_pool.generation;
    const job_t jb = g_pool.job;
    pthread_mutex_unlock(&g_pool.mu);
    run_job(&jb, wk);
    pthread_mutex_lock(&g_pool.mu);
    if (--g_pool.active == 0) pthread_cond_signal(&g_pool.done);
    pthread_mutex_unlock(&g_pool.mu);
  }
  free(wk);
  return NULL;
}

/* threads <= 1: the calling thread only.  The pool is created on first use and resized when `threads` changes. */
int fqc_solve_multi(int N, int force_final, int n_prob, const double* x0, const double* xf, const double* lim, const int* poly_ofs,
                    const int* face_ofs, const double* Ab, const int* cand_ofs, const double* dt, const uint8_t* sigma,
                    uint8_t* feasible, double* cost, double* coeffs, int32_t* iters, int threads)
{
  const plan_t* pl = get_plan(N, force_final);
  if (!pl || n_prob <= 0) return -1;
  for (int j = 0; j < n_prob; j++)
    if (poly_ofs[j + 1] - poly_ofs[j] > 63) return -1;
  job_t jb = { pl, n_prob, x0, xf, lim, Ab, dt, poly_ofs, face_ofs, cand_ofs, sigma, feasible, cost, coeffs, iters, cand_ofs[n_prob], NULL };
  memo_t* memo = NULL;
  if (g_memo_on && cand_ofs[n_prob] >= 4L * n_prob) { memo = (memo_t*)calloc((size_t)n_prob, sizeof(memo_t)); jb.memo = memo; }
  atomic_store(&g_pool.next, 0);
  if (threads > 512) threads = 512;
  if (threads <= 1)
  {
    work_t* wk = (work_t*)malloc(sizeof(work_t));
    run_job(&jb, wk);
    free(wk);
    free(memo);
    return 0;
  }
  pthread_mutex_lock(&g_pool.mu);
  if (g_pool.started && g_pool.n_threads != threads - 1)
  { /* resize: stop the old workers */
    g_pool.quit = 1;
    pthread_cond_broadcast(&g_pool.go);
    pthread_mutex_unlock(&g_pool.mu);
    for (int i = 0; i < g_pool.n_threads; i++) pthread_join(g_pool.th[i], NULL);
    pthread_mutex_lock(&g_pool.mu);
    g_pool.quit = 0; g_pool.started = 0;
  }
  if (!g_pool.started)
  {
    g_pool.n_threads = threads - 1;
    for (int i = 0; i < g_pool.n_threads; i++) pthread_create(&g_pool.th[i], NULL, worker, NULL);
    g_pool.started = 1;
  }
  g_pool.job = jb;
  g_pool.active = g_pool.n_threads;
  g_pool.generation++;
  pthread_cond_broadcast(&g_pool.go);
  pthread_mutex_unlock(&g_pool.mu);
  {
    work_t* wk = (work_t*)malloc(sizeof(work_t));
    run_job(&jb, wk);                       /* the caller works too */
    free(wk);
  }
  pthread_mutex_lock(&g_pool.mu);
  while (g_pool.active > 0) pthread_cond_wait(&g_pool.done, &g_pool.mu);
  pthread_mutex_unlock(&g_pool.mu);
  free(memo);
  return 0;
}

int fqc_abi_version(void) { return 1; }

/* ---- the chained replan on the CPU (whole sweep -> selection -> R -> safe sweep -> selection), all in C so that the CPU
 *      arm of bench.py is not slowed by interpreter overhead between the sweeps.  Same steps as oracle/pair_oracle.py and
 *      as the device chain (faster.cpp:406-430,:474-475,:521-537); getDTInitial / resetX / fillX come from fq_oracle.c. */
double fqo_dt_initial(const double* x0, const double* xf, const double* lim, int N);
int fqo_num_samples(int N, double dt, double DC);
void fqo_fill_x(int N, const double* coeffs, double dt, double DC, int n_samples, double* out);

typedef struct
{
  int whole_dt_index, whole_sigma_index, safe_dt_index, safe_sigma_index;
  double whole_cost, safe_cost, whole_dt, safe_dt, whole_dt_base, safe_dt_base;
  int n_samples_whole, k_safe;
  double R[9];
} fqc_pair_result;            /* same layout as fq_pair_result of include/faster_b200.h */

static void select_winner(const uint8_t* f, const double* c, int nf, int ns, int* di, int* si)
{
  *di = -1; *si = -1;
  for (int d = 0; d < nf && *di < 0; d++)
  {
    double best = INFINITY;
    for (int s = 0; s < ns; s++)
      if (f[d * ns + s] && c[d * ns + s] < best) { best = c[d * ns + s]; *di = d; *si = s; }
  }
}

int fqc_replan_pairs(int n_prob, int Nw, int Ns, double DC, double r_fraction, const double* x0, const double* xf_whole,
                     const double* xf_safe, const double* lim, const int* poly_ofs_w, const int* face_ofs_w, const double* Ab_w,
                     const int* poly_ofs_s, const int* face_ofs_s, const double* Ab_s, int n_fac_w, const double* fac_w, int n_sig_w,
                     const uint8_t* sig_w, int n_fac_s, const double* fac_s, int n_sig_s, const uint8_t* sig_s, uint8_t* feas_w,
                     double* cost_w, uint8_t* feas_s, double* cost_s, double* coeffs_w, double* coeffs_s, fqc_pair_result* res,
                     int threads)
{
  const long per_w = (long)n_fac_w * n_sig_w, per_s = (long)n_fac_s * n_sig_s;
  const long ncw = per_w * n_prob, ncs = per_s * n_prob;
  const int Nmax = Nw > Ns ? Nw : Ns;
  double* dt = (double*)malloc(sizeof(double) * (size_t)(ncw > ncs ? ncw : ncs));
  uint8_t* sg = (uint8_t*)malloc((size_t)(ncw > ncs ? ncw : ncs) * Nmax);
  int* co = (int*)malloc(sizeof(int) * (size_t)(n_prob + 1));
  double* x0s = (double*)malloc(sizeof(double) * 9 * (size_t)n_prob);
  double* wdt = (double*)malloc(sizeof(double) * (size_t)n_prob);
  uint8_t* wsg = (uint8_t*)malloc((size_t)n_prob * Nmax);
  uint8_t* wf = (uint8_t*)malloc((size_t)n_prob);
  double* wc = (double*)malloc(sizeof(double) * (size_t)n_prob);
  double* X = NULL;
  size_t Xcap = 0;
  int rc = 0;
  /* whole sweep */
  for (int j = 0; j < n_prob; j++)
  {
    const double b = fmax(fqo_dt_initial(x0 + 9 * j, xf_whole + 9 * j, lim + 3 * j, Nw), 2 * DC);
    res[j].whole_dt_base = b;
    for (int f = 0; f < n_fac_w; f++)
      for (int s = 0; s < n_sig_w; s++)
      {
        const long c = j * per_w + (long)f * n_sig_w + s;
        dt[c] = fac_w[f] * b;
        memcpy(sg + c * Nw, sig_w + (size_t)s * Nw, (size_t)Nw);
      }
    co[j] = (int)(j * per_w);
  }
  co[n_prob] = (int)ncw;
  rc |= fqc_solve_multi(Nw, 1, n_prob, x0, xf_whole, lim, poly_ofs_w, face_ofs_w, Ab_w, co, dt, sg, feas_w, cost_w, NULL, NULL, threads);
  for (int j = 0; j < n_prob; j++)
  {
    int di, si;
    select_winner(feas_w + j * per_w, cost_w + j * per_w, n_fac_w, n_sig_w, &di, &si);
    res[j].whole_dt_index = di; res[j].whole_sigma_index = si;
    wdt[j] = di >= 0 ? fac_w[di] * res[j].whole_dt_base : 1.0;
    memcpy(wsg + (size_t)j * Nw, sig_w + (size_t)(si >= 0 ? si : 0) * Nw, (size_t)Nw);
    co[j] = j;
  }
  co[n_prob] = n_prob;
  rc |= fqc_solve_multi(Nw, 1, n_prob, x0, xf_whole, lim, poly_ofs_w, face_ofs_w, Ab_w, co, wdt, wsg, wf, wc, coeffs_w, NULL, threads);
  /* R and the safe sweep from it */
  for (int j = 0; j < n_prob; j++)
  {
    fqc_pair_result* r = &res[j];
    if (r->whole_dt_index < 0)
    {
      r->whole_cost = INFINITY; r->whole_dt = NAN; r->n_samples_whole = 0; r->k_safe = -1; r->safe_dt_base = NAN;
      for (int i = 0; i < 9; i++) { r->R[i] = NAN; x0s[9 * j + i] = NAN; }
      continue;
    }
    r->whole_cost = wc[j]; r->whole_dt = wdt[j];
    const int n = fqo_num_samples(Nw, wdt[j], DC);
    int k = (int)(r_fraction * (double)n);
    k = k < 0 ? 0 : (k > n - 1 ? n - 1 : k);
    if ((size_t)n * 12 > Xcap) { Xcap = (size_t)n * 12 + 1200; X = (double*)realloc(X, sizeof(double) * Xcap); }
    fqo_fill_x(Nw, coeffs_w + (size_t)j * 12 * Nw, wdt[j], DC, n, X);
    for (int i = 0; i < 9; i++) { r->R[i] = X[(size_t)k * 12 + i]; x0s[9 * j + i] = r->R[i]; }
    r->n_samples_whole = n; r->k_safe = k;
    r->safe_dt_base = fmax(fqo_dt_initial(r->R, xf_safe + 9 * j, lim + 3 * j, Ns), 2 * DC);
  }
  for (int j = 0; j < n_prob; j++)
  {
    const double b = res[j].safe_dt_base;          /* NaN: the candidates of this corridor are "not solved" */
    for (int f = 0; f < n_fac_s; f++)
      for (int s = 0; s < n_sig_s; s++)
      {
        const long c = j * per_s + (long)f * n_sig_s + s;
        dt[c] = fac_s[f] * b;
        memcpy(sg + c * Ns, sig_s + (size_t)s * Ns, (size_t)Ns);
      }
    co[j] = (int)(j * per_s);
  }
  co[n_prob] = (int)ncs;
  rc |= fqc_solve_multi(Ns, 0, n_prob, x0s, xf_safe, lim, poly_ofs_s, face_ofs_s, Ab_s, co, dt, sg, feas_s, cost_s, NULL, NULL, threads);
  for (int j = 0; j < n_prob; j++)
  {
    int di, si;
    select_winner(feas_s + j * per_s, cost_s + j * per_s, n_fac_s, n_sig_s, &di, &si);
    res[j].safe_dt_index = di; res[j].safe_sigma_index = si;
    wdt[j] = di >= 0 ? fac_s[di] * res[j].safe_dt_base : NAN;
    memcpy(wsg + (size_t)j * Ns, sig_s + (size_t)(si >= 0 ? si : 0) * Ns, (size_t)Ns);
    co[j] = j;
  }
  co[n_prob] = n_prob;
  if (coeffs_s)
    rc |= fqc_solve_multi(Ns, 0, n_prob, x0s, xf_safe, lim, poly_ofs_s, face_ofs_s, Ab_s, co, wdt, wsg, wf, wc, coeffs_s, NULL, threads);
  for (int j = 0; j < n_prob; j++)
  {
    const int di = res[j].safe_dt_index;
    res[j].safe_dt = di >= 0 ? wdt[j] : NAN;
    res[j].safe_cost = di >= 0 ? cost_s[j * per_s + (long)di * n_sig_s + res[j].safe_sigma_index] : INFINITY;
  }
  free(dt); free(sg); free(co); free(x0s); free(wdt); free(wsg); free(wf); free(wc); free(X);
  return rc;
}

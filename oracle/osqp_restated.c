/* ORACLE (test infrastructure, not product): CPU restatement of the OSQP 0.6.2 algorithm.
 *
 * The reference solves its MPC QP through OsqpEigen::Solver -> OSQP 0.6.2
 * (mpcPlanner.cpp:436-527; third_party/osqp/constants.h:12 gives the version).  OSQP's source is a
 * third-party dependency that is NOT in /root/reference (only headers + lib/x86/libosqp.so), so this
 * file restates the published OSQP 0.6.2 algorithm (Stellato et al., "OSQP: an operator splitting
 * solver for quadratic programs", and the 0.6.2 release's auxil/scaling/proj/kkt step structure whose
 * prototypes ARE in the tree: third_party/osqp/auxil.h:21-154, scaling.h, proj.h, kkt.h:15-18), with
 * every constant taken from third_party/osqp/constants.h:59-118.  It is pinned against the reference's
 * binary by tests/test_oracle.py (same status / iterations / rho updates, x and obj to ~1e-9).
 *
 * Generic CSC in, FP64, c_int = long long.  The KKT system [P+sigma I, A'; A, -diag(1/rho)]
 * (kkt.h:15-18) is factored with a plain up-looking sparse LDL' in natural order (the binary uses
 * QDLDL after an AMD permutation; the factorisation is mathematically the same, rounding differs at
 * the 1e-13 level).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use this.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef long long c_int;
typedef double c_float;

/* constants.h:59-118 */
#define RHO_MIN 1e-06
#define RHO_MAX 1e06
#define RHO_EQ_OVER_RHO_INEQ 1e03
#define RHO_TOL 1e-04
#define MIN_SCALING 1e-04
#define MAX_SCALING 1e+04
#define OSQP_INFTY 1e30
#define OSQP_NAN ((c_float)0x7fc00000UL) /* constants.h:95-97: the NUMBER 2143289344.0, not a NaN */

enum { ST_DUAL_INF_INACC = 4, ST_PRIM_INF_INACC = 3, ST_SOLVED_INACC = 2, ST_SOLVED = 1, ST_MAX_ITER = -2,
       ST_PRIM_INF = -3, ST_DUAL_INF = -4, ST_NON_CVX = -7, ST_UNSOLVED = -10 };

typedef struct {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, time_limit, adaptive_rho_tolerance;
  long long max_iter, adaptive_rho, adaptive_rho_interval, check_termination, scaling, warm_start, scaled_termination;
} overrides_t;   /* same layout as ref_driver.c's ref_overrides */

typedef struct {
  c_float rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
  c_int max_iter, adaptive_rho, adaptive_rho_interval, check_termination, scaling, scaled_termination;
} settings_t;

static void default_settings(settings_t *s) { /* constants.h:59-118 (osqp_set_default_settings) */
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6; s->eps_abs = 1e-3; s->eps_rel = 1e-3;
  s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4; s->adaptive_rho_tolerance = 5;
  s->max_iter = 4000; s->adaptive_rho = 1; s->adaptive_rho_interval = 0; s->check_termination = 25;
  s->scaling = 10; s->scaled_termination = 0;
}

typedef struct {
  c_int n, m;
  /* scaled problem data (copies) */
  c_int *Pp, *Pi; c_float *Px; c_int nnzP;
  c_int *Ap, *Ai; c_float *Ax; c_int nnzA;
  c_float *q, *l, *u;
  /* scaling */
  c_float c, cinv, *D, *Dinv, *E, *Einv;
  /* rho */
  c_float *rho_vec, *rho_inv_vec; c_int *constr_type;
  /* iterates */
  c_float *x, *y, *z, *xz_tilde, *x_prev, *z_prev, *Axv, *Pxv, *Aty, *delta_y, *Atdelta_y, *delta_x, *Pdelta_x, *Adelta_x;
  /* KKT + LDL */
  c_int kn; c_int *Kp, *Ki; c_float *Kx; c_int *rho_pos; /* index of -1/rho diagonal entries in Kx */
  const c_int *perm; /* optional fill-reducing ordering of the KKT (new index -> old index), shared by a batch */
  c_int *etree, *Lnz, *Lp, *Li; c_float *Lx, *Dl, *Dinvl; c_int *iwork; unsigned char *bwork; c_float *fwork, *sol;
  settings_t s;
  /* info */
  c_int iter, status_val, rho_updates; c_float obj_val, pri_res, dua_res;
} work_t;

/* ---------- small linear algebra (lin_alg.h semantics) ---------- */
static c_float norm_inf(const c_float *v, c_int n) { c_float m = 0; for (c_int i = 0; i < n; i++) { c_float a = fabs(v[i]); if (a > m) m = a; } return m; }
static c_float scaled_norm_inf(const c_float *S, const c_float *v, c_int n) { c_float m = 0; for (c_int i = 0; i < n; i++) { c_float a = fabs(S[i] * v[i]); if (a > m) m = a; } return m; }
static c_float dot(const c_float *a, const c_float *b, c_int n) { c_float s = 0; for (c_int i = 0; i < n; i++) s += a[i] * b[i]; return s; }
static void mat_vec(c_int n, const c_int *p, const c_int *ix, const c_float *v, const c_float *x, c_float *y, int plus_eq, c_int rows) {
  if (!plus_eq) for (c_int i = 0; i < rows; i++) y[i] = 0;
  for (c_int j = 0; j < n; j++) for (c_int k = p[j]; k < p[j + 1]; k++) y[ix[k]] += v[k] * x[j];
}
static void mat_tpose_vec(c_int n, const c_int *p, const c_int *ix, const c_float *v, const c_float *x, c_float *y, int plus_eq, int skip_diag) {
  if (!plus_eq) for (c_int j = 0; j < n; j++) y[j] = 0;
  for (c_int j = 0; j < n; j++) for (c_int k = p[j]; k < p[j + 1]; k++) { if (skip_diag && ix[k] == j) continue; y[j] += v[k] * x[ix[k]]; }
}
static void sym_mat_vec(const work_t *w, const c_float *x, c_float *y) { /* P upper-triangular: y = P x */
  mat_vec(w->n, w->Pp, w->Pi, w->Px, x, y, 0, w->n);
  mat_tpose_vec(w->n, w->Pp, w->Pi, w->Px, x, y, 1, 1);
}
static c_float quad_form(const work_t *w, const c_float *x) {
  c_float qf = 0;
  for (c_int j = 0; j < w->n; j++) for (c_int k = w->Pp[j]; k < w->Pp[j + 1]; k++) {
    c_int i = w->Pi[k];
    if (i == j) qf += .5 * w->Px[k] * x[i] * x[i]; else if (i < j) qf += w->Px[k] * x[i] * x[j];
  }
  return qf;
}

/* ---------- scaling (scaling.h: scale_data, limit_scaling) ---------- */
static void limit_scaling(c_float *D, c_int n) {
  for (c_int i = 0; i < n; i++) { D[i] = D[i] < MIN_SCALING ? 1.0 : D[i]; D[i] = D[i] > MAX_SCALING ? MAX_SCALING : D[i]; }
}
static void inf_norm_cols_sym_triu(const work_t *w, c_float *E) {
  for (c_int j = 0; j < w->n; j++) E[j] = 0;
  for (c_int j = 0; j < w->n; j++) for (c_int k = w->Pp[j]; k < w->Pp[j + 1]; k++) {
    c_int i = w->Pi[k]; c_float a = fabs(w->Px[k]);
    if (a > E[j]) E[j] = a;
    if (i != j && a > E[i]) E[i] = a;
  }
}
static void scale_data(work_t *w) {
  c_int n = w->n, m = w->m;
  c_float *Dt = (c_float *)malloc(sizeof(c_float) * (size_t)(n + 1)), *DtA = (c_float *)malloc(sizeof(c_float) * (size_t)(n + 1)), *Et = (c_float *)malloc(sizeof(c_float) * (size_t)(m + 1));
  w->c = 1.0;
  for (c_int i = 0; i < n; i++) w->D[i] = w->Dinv[i] = 1.;
  for (c_int i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1.;
  for (c_int it = 0; it < w->s.scaling; it++) {
    /* column inf-norms of [P A'; A 0] */
    inf_norm_cols_sym_triu(w, Dt);
    for (c_int j = 0; j < n; j++) { DtA[j] = 0; for (c_int k = w->Ap[j]; k < w->Ap[j + 1]; k++) { c_float a = fabs(w->Ax[k]); if (a > DtA[j]) DtA[j] = a; } }
    for (c_int j = 0; j < n; j++) if (DtA[j] > Dt[j]) Dt[j] = DtA[j];
    for (c_int i = 0; i < m; i++) Et[i] = 0;
    for (c_int j = 0; j < n; j++) for (c_int k = w->Ap[j]; k < w->Ap[j + 1]; k++) { c_float a = fabs(w->Ax[k]); if (a > Et[w->Ai[k]]) Et[w->Ai[k]] = a; }
    limit_scaling(Dt, n); limit_scaling(Et, m);
    for (c_int j = 0; j < n; j++) Dt[j] = 1. / sqrt(Dt[j]);
    for (c_int i = 0; i < m; i++) Et[i] = 1. / sqrt(Et[i]);
    /* P <- D P D (pre then post), A <- E A D, q <- D q */
    for (c_int j = 0; j < n; j++) for (c_int k = w->Pp[j]; k < w->Pp[j + 1]; k++) w->Px[k] *= Dt[w->Pi[k]];
    for (c_int j = 0; j < n; j++) for (c_int k = w->Pp[j]; k < w->Pp[j + 1]; k++) w->Px[k] *= Dt[j];
    for (c_int j = 0; j < n; j++) for (c_int k = w->Ap[j]; k < w->Ap[j + 1]; k++) w->Ax[k] *= Et[w->Ai[k]];
    for (c_int j = 0; j < n; j++) for (c_int k = w->Ap[j]; k < w->Ap[j + 1]; k++) w->Ax[k] *= Dt[j];
    for (c_int j = 0; j < n; j++) w->q[j] *= Dt[j];
    for (c_int j = 0; j < n; j++) w->D[j] *= Dt[j];
    for (c_int i = 0; i < m; i++) w->E[i] *= Et[i];
    /* cost normalisation */
    inf_norm_cols_sym_triu(w, Dt);
    c_float c_temp = 0; for (c_int j = 0; j < n; j++) c_temp += Dt[j]; c_temp /= (c_float)n;
    c_float nq = norm_inf(w->q, n);
    limit_scaling(&nq, 1);
    if (nq > c_temp) c_temp = nq;
    limit_scaling(&c_temp, 1);
    c_temp = 1. / c_temp;
    for (c_int k = 0; k < w->nnzP; k++) w->Px[k] *= c_temp;
    for (c_int j = 0; j < n; j++) w->q[j] *= c_temp;
    w->c *= c_temp;
  }
  w->cinv = 1. / w->c;
  for (c_int j = 0; j < n; j++) w->Dinv[j] = 1. / w->D[j];
  for (c_int i = 0; i < m; i++) w->Einv[i] = 1. / w->E[i];
  for (c_int i = 0; i < m; i++) { w->l[i] *= w->E[i]; w->u[i] *= w->E[i]; }
  free(Dt); free(DtA); free(Et);
}

/* ---------- rho vector (auxil.h: set_rho_vec / update_rho_vec) ---------- */
static void set_rho_vec(work_t *w) {
  w->s.rho = fmin(fmax(w->s.rho, RHO_MIN), RHO_MAX);
  for (c_int i = 0; i < w->m; i++) {
    if ((w->l[i] < -OSQP_INFTY * MIN_SCALING) && (w->u[i] > OSQP_INFTY * MIN_SCALING)) { w->constr_type[i] = -1; w->rho_vec[i] = RHO_MIN; }
    else if (w->u[i] - w->l[i] < RHO_TOL) { w->constr_type[i] = 1; w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->s.rho; }
    else { w->constr_type[i] = 0; w->rho_vec[i] = w->s.rho; }
    w->rho_inv_vec[i] = 1. / w->rho_vec[i];
  }
}

/* ---------- sparse LDL' (natural order), Davis-style up-looking ---------- */
static int ldl_symbolic(work_t *w) {
  c_int n = w->kn; const c_int *Ap = w->Kp, *Ai = w->Ki;
  c_int *etree = w->etree, *Lnz = w->Lnz, *flag = w->iwork;
  for (c_int k = 0; k < n; k++) {
    etree[k] = -1; flag[k] = k; Lnz[k] = 0;
    for (c_int p = Ap[k]; p < Ap[k + 1]; p++) {
      c_int i = Ai[p];
      if (i > k) return -1;
      for (; flag[i] != k; i = etree[i]) { if (etree[i] == -1) etree[i] = k; Lnz[i]++; flag[i] = k; }
    }
  }
  w->Lp[0] = 0; for (c_int k = 0; k < n; k++) w->Lp[k + 1] = w->Lp[k] + Lnz[k];
  return 0;
}
static int ldl_numeric(work_t *w) {
  c_int n = w->kn; const c_int *Ap = w->Kp, *Ai = w->Ki; const c_float *Ax = w->Kx;
  c_int *Lp = w->Lp, *Li = w->Li, *etree = w->etree; c_float *Lx = w->Lx, *D = w->Dl;
  c_int *flag = w->iwork, *pattern = w->iwork + n, *lnz = w->iwork + 2 * n; c_float *y = w->fwork;
  for (c_int k = 0; k < n; k++) {
    y[k] = 0.0; c_int top = n; flag[k] = k; lnz[k] = 0;
    for (c_int p = Ap[k]; p < Ap[k + 1]; p++) {
      c_int i = Ai[p];
      y[i] += Ax[p];
      c_int len;
      for (len = 0; flag[i] != k; i = etree[i]) { pattern[len++] = i; flag[i] = k; }
      while (len > 0) pattern[--top] = pattern[--len];
    }
    D[k] = y[k]; y[k] = 0.0;
    for (; top < n; top++) {
      c_int i = pattern[top]; c_float yi = y[i]; y[i] = 0.0;
      c_int p2 = Lp[i] + lnz[i];
      for (c_int p = Lp[i]; p < p2; p++) y[Li[p]] -= Lx[p] * yi;
      c_float l_ki = yi / D[i];
      D[k] -= l_ki * yi;
      Li[p2] = k; Lx[p2] = l_ki; lnz[i]++;
    }
    if (D[k] == 0.0) return -1;
    w->Dinvl[k] = 1.0 / D[k];
  }
  return 0;
}
static void ldl_solve(const work_t *w, c_float *x) {
  c_int n = w->kn; const c_int *Lp = w->Lp, *Li = w->Li; const c_float *Lx = w->Lx;
  for (c_int j = 0; j < n; j++) { c_float xj = x[j]; for (c_int p = Lp[j]; p < Lp[j + 1]; p++) x[Li[p]] -= Lx[p] * xj; }
  for (c_int j = 0; j < n; j++) x[j] *= w->Dinvl[j];
  for (c_int j = n - 1; j >= 0; j--) { c_float xj = x[j]; for (c_int p = Lp[j]; p < Lp[j + 1]; p++) xj -= Lx[p] * x[Li[p]]; x[j] = xj; }
}

/* Greedy minimum-degree ordering on the KKT graph (stand-in for the AMD step of the binary; any
   symmetric permutation is valid for a quasi-definite LDL').  perm[new] = old.  Done once per pattern. */
static c_int *kkt_min_degree(c_int n, c_int m, const c_int *Pp, const c_int *Pi, const c_int *Ap, const c_int *Ai) {
  c_int kn = n + m, W = (kn + 63) / 64;
  unsigned long long *adj = (unsigned long long *)calloc((size_t)(kn * W), 8);
  c_int *deg = (c_int *)calloc((size_t)kn, sizeof(c_int)), *perm = (c_int *)malloc(sizeof(c_int) * (size_t)kn);
  unsigned char *done = (unsigned char *)calloc((size_t)kn, 1);
#define SETB(a, b) adj[(a) * W + ((b) >> 6)] |= 1ULL << ((b) & 63)
  for (c_int j = 0; j < n; j++) {
    for (c_int k = Pp[j]; k < Pp[j + 1]; k++) if (Pi[k] != j) { SETB(j, Pi[k]); SETB(Pi[k], j); }
    for (c_int k = Ap[j]; k < Ap[j + 1]; k++) { SETB(j, n + Ai[k]); SETB(n + Ai[k], j); }
  }
  for (c_int v = 0; v < kn; v++) { c_int d = 0; for (c_int w = 0; w < W; w++) d += __builtin_popcountll(adj[v * W + w]); deg[v] = d; }
  for (c_int step = 0; step < kn; step++) {
    c_int best = -1;
    for (c_int v = 0; v < kn; v++) if (!done[v] && (best < 0 || deg[v] < deg[best])) best = v;
    perm[step] = best; done[best] = 1;
    unsigned long long *av = adj + best * W;
    for (c_int w = 0; w < W; w++) {
      unsigned long long bits = av[w];
      while (bits) {
        c_int u = w * 64 + __builtin_ctzll(bits); bits &= bits - 1;
        unsigned long long *au = adj + u * W;
        for (c_int t = 0; t < W; t++) au[t] |= av[t];
        au[best >> 6] &= ~(1ULL << (best & 63)); au[u >> 6] &= ~(1ULL << (u & 63));
      }
    }
    for (c_int w = 0; w < W; w++) {        /* neighbours lose `best`; recount their degree */
      unsigned long long bits = av[w];
      while (bits) {
        c_int u = w * 64 + __builtin_ctzll(bits); bits &= bits - 1;
        c_int d = 0; for (c_int t = 0; t < W; t++) d += __builtin_popcountll(adj[u * W + t]); deg[u] = d;
      }
    }
  }
  free(adj); free(deg); free(done);
  return perm;
}

/* KKT upper triangle in CSC: [P+sigma I, A'; A, -diag(1/rho)] (kkt.h:15-18), symmetrically permuted by w->perm */
static int form_kkt(work_t *w) {
  c_int n = w->n, m = w->m, kn = n + m;
  w->kn = kn;
  c_int nnz = w->nnzP + n + w->nnzA + m;
  c_int *ti = (c_int *)malloc(sizeof(c_int) * (size_t)nnz), *tj = (c_int *)malloc(sizeof(c_int) * (size_t)nnz); c_float *tv = (c_float *)malloc(sizeof(c_float) * (size_t)nnz);
  c_int *tag = (c_int *)malloc(sizeof(c_int) * (size_t)nnz);   /* >=0: rho row index for the -1/rho entries */
  c_int *inv = (c_int *)malloc(sizeof(c_int) * (size_t)kn);
  for (c_int k = 0; k < kn; k++) inv[w->perm ? w->perm[k] : k] = k;
  c_int cnt = 0;
  for (c_int j = 0; j < n; j++) {
    int has_diag = 0;
    for (c_int k = w->Pp[j]; k < w->Pp[j + 1]; k++) {
      if (w->Pi[k] > j) return -1;
      ti[cnt] = w->Pi[k]; tj[cnt] = j; tv[cnt] = w->Px[k]; tag[cnt] = -1;
      if (w->Pi[k] == j) { tv[cnt] += w->s.sigma; has_diag = 1; }
      cnt++;
    }
    if (!has_diag) { ti[cnt] = j; tj[cnt] = j; tv[cnt] = w->s.sigma; tag[cnt] = -1; cnt++; }
    for (c_int k = w->Ap[j]; k < w->Ap[j + 1]; k++) { ti[cnt] = j; tj[cnt] = n + w->Ai[k]; tv[cnt] = w->Ax[k]; tag[cnt] = -1; cnt++; }
  }
  for (c_int i = 0; i < m; i++) { ti[cnt] = n + i; tj[cnt] = n + i; tv[cnt] = -w->rho_inv_vec[i]; tag[cnt] = i; cnt++; }
  /* permute, force upper triangle, bucket by column */
  w->Kp = (c_int *)calloc((size_t)kn + 1, sizeof(c_int)); w->Ki = (c_int *)malloc(sizeof(c_int) * (size_t)cnt); w->Kx = (c_float *)malloc(sizeof(c_float) * (size_t)cnt);
  w->rho_pos = (c_int *)malloc(sizeof(c_int) * (size_t)(m + 1));
  for (c_int e = 0; e < cnt; e++) { c_int a = inv[ti[e]], b = inv[tj[e]]; if (a > b) { c_int t = a; a = b; b = t; } ti[e] = a; tj[e] = b; w->Kp[b + 1]++; }
  for (c_int k = 0; k < kn; k++) w->Kp[k + 1] += w->Kp[k];
  c_int *fill = (c_int *)calloc((size_t)kn, sizeof(c_int));
  for (c_int e = 0; e < cnt; e++) { c_int d = w->Kp[tj[e]] + fill[tj[e]]++; w->Ki[d] = ti[e]; w->Kx[d] = tv[e]; if (tag[e] >= 0) w->rho_pos[tag[e]] = d; }
  free(ti); free(tj); free(tv); free(tag); free(fill); free(inv);
  w->etree = (c_int *)malloc(sizeof(c_int) * (size_t)kn); w->Lnz = (c_int *)malloc(sizeof(c_int) * (size_t)kn); w->Lp = (c_int *)malloc(sizeof(c_int) * (size_t)(kn + 1));
  w->iwork = (c_int *)malloc(sizeof(c_int) * (size_t)(3 * kn)); w->fwork = (c_float *)malloc(sizeof(c_float) * (size_t)kn); w->sol = (c_float *)malloc(sizeof(c_float) * (size_t)kn);
  w->Dl = (c_float *)malloc(sizeof(c_float) * (size_t)kn); w->Dinvl = (c_float *)malloc(sizeof(c_float) * (size_t)kn);
  if (ldl_symbolic(w)) return -1;
  c_int lnz = w->Lp[kn];
  w->Li = (c_int *)malloc(sizeof(c_int) * (size_t)(lnz + 1)); w->Lx = (c_float *)malloc(sizeof(c_float) * (size_t)(lnz + 1));
  return ldl_numeric(w);
}
static int refactor_rho(work_t *w) {
  for (c_int i = 0; i < w->m; i++) w->Kx[w->rho_pos[i]] = -w->rho_inv_vec[i];
  return ldl_numeric(w);
}

/* ---------- ADMM steps (auxil.h:67-112) ---------- */
static void update_xz_tilde(work_t *w) {
  c_int n = w->n, m = w->m;
  for (c_int i = 0; i < n; i++) w->xz_tilde[i] = w->s.sigma * w->x_prev[i] - w->q[i];
  for (c_int i = 0; i < m; i++) w->xz_tilde[i + n] = w->z_prev[i] - w->rho_inv_vec[i] * w->y[i];
  if (w->perm) { for (c_int k = 0; k < n + m; k++) w->fwork[k] = w->xz_tilde[w->perm[k]]; ldl_solve(w, w->fwork); for (c_int k = 0; k < n + m; k++) w->sol[w->perm[k]] = w->fwork[k]; }
  else { memcpy(w->sol, w->xz_tilde, sizeof(c_float) * (size_t)(n + m)); ldl_solve(w, w->sol); }
  for (c_int j = 0; j < n; j++) w->xz_tilde[j] = w->sol[j];
  for (c_int j = 0; j < m; j++) w->xz_tilde[j + n] += w->rho_inv_vec[j] * w->sol[j + n];
}
static void update_x(work_t *w) {
  for (c_int i = 0; i < w->n; i++) w->x[i] = w->s.alpha * w->xz_tilde[i] + (1.0 - w->s.alpha) * w->x_prev[i];
  for (c_int i = 0; i < w->n; i++) w->delta_x[i] = w->x[i] - w->x_prev[i];
}
static void update_z(work_t *w) {
  c_int n = w->n;
  for (c_int i = 0; i < w->m; i++) w->z[i] = w->s.alpha * w->xz_tilde[i + n] + (1.0 - w->s.alpha) * w->z_prev[i] + w->rho_inv_vec[i] * w->y[i];
  for (c_int i = 0; i < w->m; i++) w->z[i] = fmin(fmax(w->z[i], w->l[i]), w->u[i]);   /* proj.h: project */
}
static void update_y(work_t *w) {
  c_int n = w->n;
  for (c_int i = 0; i < w->m; i++) {
    w->delta_y[i] = w->rho_vec[i] * (w->s.alpha * w->xz_tilde[i + n] + (1.0 - w->s.alpha) * w->z_prev[i] - w->z[i]);
    w->y[i] += w->delta_y[i];
  }
}
static c_float compute_obj_val(const work_t *w, const c_float *x) { return (quad_form(w, x) + dot(w->q, x, w->n)) * w->cinv; }
static c_float compute_pri_res(work_t *w) {   /* z_prev is scratch: Ax - z (scaled) */
  mat_vec(w->n, w->Ap, w->Ai, w->Ax, w->x, w->Axv, 0, w->m);
  for (c_int i = 0; i < w->m; i++) w->z_prev[i] = w->Axv[i] - w->z[i];
  if (w->s.scaling && !w->s.scaled_termination) return scaled_norm_inf(w->Einv, w->z_prev, w->m);
  return norm_inf(w->z_prev, w->m);
}
static c_float compute_dua_res(work_t *w) {   /* x_prev is scratch: q + Px + A'y (scaled) */
  memcpy(w->x_prev, w->q, sizeof(c_float) * (size_t)w->n);
  sym_mat_vec(w, w->x, w->Pxv);
  for (c_int i = 0; i < w->n; i++) w->x_prev[i] += w->Pxv[i];
  if (w->m > 0) { mat_tpose_vec(w->n, w->Ap, w->Ai, w->Ax, w->y, w->Aty, 0, 0); for (c_int i = 0; i < w->n; i++) w->x_prev[i] += w->Aty[i]; }
  if (w->s.scaling && !w->s.scaled_termination) return w->cinv * scaled_norm_inf(w->Dinv, w->x_prev, w->n);
  return norm_inf(w->x_prev, w->n);
}
static void update_info(work_t *w, c_int iter) {
  w->pri_res = w->m == 0 ? 0. : compute_pri_res(w);
  w->dua_res = compute_dua_res(w);
  w->iter = iter;
}
static c_float compute_pri_tol(const work_t *w, c_float eps_abs, c_float eps_rel) {
  c_float a, b;
  if (w->s.scaling && !w->s.scaled_termination) { a = scaled_norm_inf(w->Einv, w->z, w->m); b = scaled_norm_inf(w->Einv, w->Axv, w->m); }
  else { a = norm_inf(w->z, w->m); b = norm_inf(w->Axv, w->m); }
  return eps_abs + eps_rel * fmax(a, b);
}
static c_float compute_dua_tol(const work_t *w, c_float eps_abs, c_float eps_rel) {
  c_float t;
  if (w->s.scaling && !w->s.scaled_termination) {
    t = scaled_norm_inf(w->Dinv, w->q, w->n);
    t = fmax(t, scaled_norm_inf(w->Dinv, w->Aty, w->n));
    t = fmax(t, scaled_norm_inf(w->Dinv, w->Pxv, w->n));
    t *= w->cinv;
  } else { t = norm_inf(w->q, w->n); t = fmax(t, norm_inf(w->Aty, w->n)); t = fmax(t, norm_inf(w->Pxv, w->n)); }
  return eps_abs + eps_rel * t;
}
static int is_primal_infeasible(work_t *w, c_float eps) {
  c_float norm_dy, ineq_lhs = 0.0;
  for (c_int i = 0; i < w->m; i++) {
    if (w->u[i] > OSQP_INFTY * MIN_SCALING) {
      if (w->l[i] < -OSQP_INFTY * MIN_SCALING) w->delta_y[i] = 0.0; else w->delta_y[i] = fmin(w->delta_y[i], 0.0);
    } else if (w->l[i] < -OSQP_INFTY * MIN_SCALING) w->delta_y[i] = fmax(w->delta_y[i], 0.0);
  }
  if (w->s.scaling && !w->s.scaled_termination) { for (c_int i = 0; i < w->m; i++) w->Adelta_x[i] = w->E[i] * w->delta_y[i]; norm_dy = norm_inf(w->Adelta_x, w->m); }
  else norm_dy = norm_inf(w->delta_y, w->m);
  if (norm_dy > eps) {
    /* IEEE semantics kept on purpose: u = +inf times max(dy,0) = 0 gives NaN, so with IEEE-inf bounds
       (mpcPlanner.cpp:913-914) the comparison below is false and infeasibility is never declared. */
    for (c_int i = 0; i < w->m; i++) ineq_lhs += w->u[i] * fmax(w->delta_y[i], 0) + w->l[i] * fmin(w->delta_y[i], 0);
    if (ineq_lhs < -eps * norm_dy) {
      mat_tpose_vec(w->n, w->Ap, w->Ai, w->Ax, w->delta_y, w->Atdelta_y, 0, 0);
      if (w->s.scaling && !w->s.scaled_termination) for (c_int i = 0; i < w->n; i++) w->Atdelta_y[i] *= w->Dinv[i];
      return norm_inf(w->Atdelta_y, w->n) < eps * norm_dy;
    }
  }
  return 0;
}
static int is_dual_infeasible(work_t *w, c_float eps) {
  c_float norm_dx, cost_scaling;
  if (w->s.scaling && !w->s.scaled_termination) { norm_dx = scaled_norm_inf(w->D, w->delta_x, w->n); cost_scaling = w->c; }
  else { norm_dx = norm_inf(w->delta_x, w->n); cost_scaling = 1.0; }
  if (norm_dx > eps) {
    if (dot(w->q, w->delta_x, w->n) < -cost_scaling * eps * norm_dx) {
      sym_mat_vec(w, w->delta_x, w->Pdelta_x);
      if (w->s.scaling && !w->s.scaled_termination) for (c_int i = 0; i < w->n; i++) w->Pdelta_x[i] *= w->Dinv[i];
      if (norm_inf(w->Pdelta_x, w->n) < cost_scaling * eps * norm_dx) {
        mat_vec(w->n, w->Ap, w->Ai, w->Ax, w->delta_x, w->Adelta_x, 0, w->m);
        if (w->s.scaling && !w->s.scaled_termination) for (c_int i = 0; i < w->m; i++) w->Adelta_x[i] *= w->Einv[i];
        for (c_int i = 0; i < w->m; i++) {
          if (((w->u[i] < OSQP_INFTY * MIN_SCALING) && (w->Adelta_x[i] > eps * norm_dx)) ||
              ((w->l[i] > -OSQP_INFTY * MIN_SCALING) && (w->Adelta_x[i] < -eps * norm_dx))) return 0;
        }
        return 1;
      }
    }
  }
  return 0;
}
static int check_termination(work_t *w, int approximate) {
  c_float eps_abs = w->s.eps_abs, eps_rel = w->s.eps_rel, epi = w->s.eps_prim_inf, edi = w->s.eps_dual_inf;
  int prim_res_check = 0, dual_res_check = 0, prim_inf_check = 0, dual_inf_check = 0;
  if ((w->pri_res > OSQP_INFTY) || (w->dua_res > OSQP_INFTY)) { w->status_val = ST_NON_CVX; w->obj_val = OSQP_NAN; return 1; }
  if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
  if (w->m == 0) prim_res_check = 1;
  else {
    c_float eps_prim = compute_pri_tol(w, eps_abs, eps_rel);
    if (w->pri_res < eps_prim) prim_res_check = 1; else prim_inf_check = is_primal_infeasible(w, epi);
  }
  c_float eps_dual = compute_dua_tol(w, eps_abs, eps_rel);
  if (w->dua_res < eps_dual) dual_res_check = 1; else dual_inf_check = is_dual_infeasible(w, edi);
  if (prim_res_check && dual_res_check) { w->status_val = approximate ? ST_SOLVED_INACC : ST_SOLVED; return 1; }
  else if (prim_inf_check) {
    w->status_val = approximate ? ST_PRIM_INF_INACC : ST_PRIM_INF;
    if (w->s.scaling && !w->s.scaled_termination) for (c_int i = 0; i < w->m; i++) w->delta_y[i] *= w->E[i];
    w->obj_val = OSQP_INFTY; return 1;
  } else if (dual_inf_check) {
    w->status_val = approximate ? ST_DUAL_INF_INACC : ST_DUAL_INF;
    if (w->s.scaling && !w->s.scaled_termination) for (c_int i = 0; i < w->n; i++) w->delta_x[i] *= w->D[i];
    w->obj_val = -OSQP_INFTY; return 1;
  }
  return 0;
}
static c_float compute_rho_estimate(const work_t *w) {   /* scaled residuals left in z_prev / x_prev by update_info */
  c_float pri_res = norm_inf(w->z_prev, w->m), dua_res = norm_inf(w->x_prev, w->n);
  c_float pn = fmax(norm_inf(w->z, w->m), norm_inf(w->Axv, w->m));
  pri_res /= (pn + 1e-10);
  c_float dn = norm_inf(w->q, w->n); dn = fmax(dn, norm_inf(w->Aty, w->n)); dn = fmax(dn, norm_inf(w->Pxv, w->n));
  dua_res /= (dn + 1e-10);
  c_float r = w->s.rho * sqrt(pri_res / (dua_res + 1e-10));
  return fmin(fmax(r, RHO_MIN), RHO_MAX);
}
static int adapt_rho(work_t *w) {
  c_float rho_new = compute_rho_estimate(w);
  if ((rho_new > w->s.rho * w->s.adaptive_rho_tolerance) || (rho_new < w->s.rho / w->s.adaptive_rho_tolerance)) {
    w->s.rho = fmin(fmax(rho_new, RHO_MIN), RHO_MAX);
    for (c_int i = 0; i < w->m; i++) {
      if (w->constr_type[i] == 0) { w->rho_vec[i] = w->s.rho; w->rho_inv_vec[i] = 1. / w->s.rho; }
      else if (w->constr_type[i] == 1) { w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->s.rho; w->rho_inv_vec[i] = 1. / w->rho_vec[i]; }
    }
    if (refactor_rho(w)) return 1;
    w->rho_updates += 1;
  }
  return 0;
}
static int has_solution(const work_t *w) {
  return (w->status_val != ST_PRIM_INF) && (w->status_val != ST_PRIM_INF_INACC) && (w->status_val != ST_DUAL_INF) &&
         (w->status_val != ST_DUAL_INF_INACC) && (w->status_val != ST_NON_CVX);
}

static void *xmalloc(size_t n) { void *p = calloc(n ? n : 1, 1); return p; }

static void free_work(work_t *w) {
  void *ptrs[] = { w->Pp, w->Pi, w->Px, w->Ap, w->Ai, w->Ax, w->q, w->l, w->u, w->D, w->Dinv, w->E, w->Einv, w->rho_vec, w->rho_inv_vec,
                   w->constr_type, w->x, w->y, w->z, w->xz_tilde, w->x_prev, w->z_prev, w->Axv, w->Pxv, w->Aty, w->delta_y, w->Atdelta_y,
                   w->delta_x, w->Pdelta_x, w->Adelta_x, w->Kp, w->Ki, w->Kx, w->rho_pos, w->etree, w->Lnz, w->Lp, w->Li, w->Lx, w->Dl,
                   w->Dinvl, w->iwork, w->fwork, w->sol };
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(ptrs[i]);
}

/* osqp_setup: copy, validate, scale, rho_vec, KKT + factor.  Returns 0 or an osqp_error_type (constants.h:41-49). */
static int port_setup(work_t *w, c_int n, c_int m, const c_int *Pp, const c_int *Pi, const c_float *Px, const c_float *q,
                      const c_int *Ap, const c_int *Ai, const c_float *Ax, const c_float *l, const c_float *u, const settings_t *s, const c_int *perm) {
  memset(w, 0, sizeof *w); w->perm = perm;
  w->n = n; w->m = m; w->s = *s; w->nnzP = Pp[n]; w->nnzA = Ap[n];
  for (c_int i = 0; i < m; i++) if (l[i] > u[i]) return 1;                     /* validate_data: l <= u */
  for (c_int j = 0; j < n; j++) for (c_int k = Pp[j]; k < Pp[j + 1]; k++) if (Pi[k] > j) return 1;  /* P upper triangular */
#define DUPI(dst, src, cnt) dst = (c_int *)xmalloc(sizeof(c_int) * (size_t)(cnt)); memcpy(dst, src, sizeof(c_int) * (size_t)(cnt))
#define DUPF(dst, src, cnt) dst = (c_float *)xmalloc(sizeof(c_float) * (size_t)(cnt)); memcpy(dst, src, sizeof(c_float) * (size_t)(cnt))
  DUPI(w->Pp, Pp, n + 1); DUPI(w->Pi, Pi, w->nnzP); DUPF(w->Px, Px, w->nnzP);
  DUPI(w->Ap, Ap, n + 1); DUPI(w->Ai, Ai, w->nnzA); DUPF(w->Ax, Ax, w->nnzA);
  DUPF(w->q, q, n); DUPF(w->l, l, m); DUPF(w->u, u, m);
#define VN(v) w->v = (c_float *)xmalloc(sizeof(c_float) * (size_t)n)
#define VM(v) w->v = (c_float *)xmalloc(sizeof(c_float) * (size_t)m)
  VN(D); VN(Dinv); VM(E); VM(Einv); VM(rho_vec); VM(rho_inv_vec); w->constr_type = (c_int *)xmalloc(sizeof(c_int) * (size_t)m);
  VN(x); VM(y); VM(z); w->xz_tilde = (c_float *)xmalloc(sizeof(c_float) * (size_t)(n + m)); VN(x_prev); VM(z_prev); VM(Axv); VN(Pxv); VN(Aty);
  VM(delta_y); VN(Atdelta_y); VN(delta_x); VN(Pdelta_x); VM(Adelta_x);
  if (w->s.scaling) scale_data(w);
  else { w->c = w->cinv = 1; for (c_int i = 0; i < n; i++) w->D[i] = w->Dinv[i] = 1; for (c_int i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1; }
  set_rho_vec(w);
  if (form_kkt(w)) return 4;
  w->status_val = ST_UNSOLVED;
  return 0;
}
static void port_warm_start(work_t *w, const c_float *x, const c_float *y) {   /* osqp.h:157 */
  for (c_int i = 0; i < w->n; i++) w->x[i] = x[i];
  for (c_int i = 0; i < w->m; i++) w->y[i] = y ? y[i] : 0.0;
  if (w->s.scaling) {
    for (c_int i = 0; i < w->n; i++) w->x[i] *= w->Dinv[i];
    for (c_int i = 0; i < w->m; i++) w->y[i] *= w->Einv[i];
    for (c_int i = 0; i < w->m; i++) w->y[i] *= w->c;
  }
  mat_vec(w->n, w->Ap, w->Ai, w->Ax, w->x, w->z, 0, w->m);
}
static int port_solve(work_t *w) {   /* osqp.h:78 (osqp_solve); no time limit, no printing, no polish */
  c_int iter; int can_check = 0; c_float *t;
  w->status_val = ST_UNSOLVED; w->rho_updates = 0;
  for (iter = 1; iter <= w->s.max_iter; iter++) {
    t = w->x; w->x = w->x_prev; w->x_prev = t;
    t = w->z; w->z = w->z_prev; w->z_prev = t;
    update_xz_tilde(w); update_x(w); update_z(w); update_y(w);
    can_check = w->s.check_termination && (iter % w->s.check_termination == 0);
    if (can_check) { update_info(w, iter); if (check_termination(w, 0)) break; }
    if (w->s.adaptive_rho && w->s.adaptive_rho_interval && (iter % w->s.adaptive_rho_interval == 0)) {
      if (!can_check) update_info(w, iter);
      if (adapt_rho(w)) return 1;
    }
  }
  if (!can_check) { update_info(w, iter - 1); check_termination(w, 0); }
  if (has_solution(w)) w->obj_val = compute_obj_val(w, w->x);
  if (w->status_val == ST_UNSOLVED) { if (!check_termination(w, 1)) w->status_val = ST_MAX_ITER; }
  return 0;
}

typedef struct {
  c_int n, m, nnzP, nnzA;
  const c_int *P_colptr, *P_rowidx, *A_colptr, *A_rowidx;
  const c_float *P_val, *q, *A_val, *l, *u, *warm_x, *warm_y;
  settings_t s;
  c_float *x, *y, *obj, *pri_res, *dua_res, *setup_time, *solve_time, *wall_time;
  c_int *status, *iter, *rho_updates, *exitflag;
  c_float *dump; c_int dump_idx; const c_int *perm;
  c_int begin, end;
} job_t;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void solve_one(const job_t *J, c_int b) {
  c_int n = J->n, m = J->m;
  double t0 = now_s();
  work_t w;
  int flag = port_setup(&w, n, m, J->P_colptr, J->P_rowidx, J->P_val + b * J->nnzP, J->q + b * n, J->A_colptr, J->A_rowidx,
                        J->A_val + b * J->nnzA, J->l + b * m, J->u + b * m, &J->s, J->perm);
  double t1 = now_s();
  if (flag) {
    J->exitflag[b] = flag; J->status[b] = ST_UNSOLVED; J->iter[b] = 0; J->rho_updates[b] = 0; J->obj[b] = 0;
    J->wall_time[b] = now_s() - t0; free_work(&w); return;
  }
  if (J->warm_x) port_warm_start(&w, J->warm_x + b * n, J->warm_y ? J->warm_y + b * m : NULL);
  int sf = port_solve(&w);
  double t2 = now_s();
  J->exitflag[b] = sf; J->status[b] = w.status_val; J->iter[b] = w.iter; J->rho_updates[b] = w.rho_updates;
  J->obj[b] = w.obj_val; J->pri_res[b] = w.pri_res; J->dua_res[b] = w.dua_res; J->setup_time[b] = t1 - t0; J->solve_time[b] = t2 - t1;
  if (J->dump && b == J->dump_idx) {
    c_float *o = J->dump; *o++ = w.c; *o++ = w.s.rho;
    memcpy(o, w.D, 8 * n); o += n; memcpy(o, w.E, 8 * m); o += m; memcpy(o, w.rho_vec, 8 * m); o += m;
    memcpy(o, w.x, 8 * n); o += n; memcpy(o, w.z, 8 * m); o += m; memcpy(o, w.y, 8 * m);
  }
  /* store_solution (auxil.h:118) */
  if (has_solution(&w)) {
    for (c_int i = 0; i < n; i++) J->x[b * n + i] = w.s.scaling ? w.D[i] * w.x[i] : w.x[i];
    if (J->y) for (c_int i = 0; i < m; i++) J->y[b * m + i] = w.s.scaling ? w.E[i] * w.y[i] * w.cinv : w.y[i];
  } else {
    for (c_int i = 0; i < n; i++) J->x[b * n + i] = OSQP_NAN;
    if (J->y) for (c_int i = 0; i < m; i++) J->y[b * m + i] = OSQP_NAN;
  }
  free_work(&w);
  J->wall_time[b] = now_s() - t0;
}
static void *worker(void *arg) { job_t *J = (job_t *)arg; for (c_int b = J->begin; b < J->end; ++b) solve_one(J, b); return NULL; }

double port_solve_batch(c_int n, c_int m, c_int nnzP, c_int nnzA, c_int B,
                        const c_int *P_colptr, const c_int *P_rowidx, const c_float *P_val, const c_float *q,
                        const c_int *A_colptr, const c_int *A_rowidx, const c_float *A_val,
                        const c_float *l, const c_float *u, const c_float *warm_x, const c_float *warm_y,
                        const overrides_t *o, int nthreads,
                        c_float *x, c_float *y, c_int *status, c_int *iter, c_int *rho_updates, c_int *exitflag,
                        c_float *obj, c_float *pri_res, c_float *dua_res, c_float *setup_time, c_float *solve_time,
                        c_float *wall_time, c_float *dump, c_int dump_idx) {
  settings_t s; default_settings(&s);
  if (o) {
    if (o->rho == o->rho) s.rho = o->rho;
    if (o->sigma == o->sigma) s.sigma = o->sigma;
    if (o->alpha == o->alpha) s.alpha = o->alpha;
    if (o->eps_abs == o->eps_abs) s.eps_abs = o->eps_abs;
    if (o->eps_rel == o->eps_rel) s.eps_rel = o->eps_rel;
    if (o->eps_prim_inf == o->eps_prim_inf) s.eps_prim_inf = o->eps_prim_inf;
    if (o->eps_dual_inf == o->eps_dual_inf) s.eps_dual_inf = o->eps_dual_inf;
    if (o->adaptive_rho_tolerance == o->adaptive_rho_tolerance) s.adaptive_rho_tolerance = o->adaptive_rho_tolerance;
    if (o->max_iter >= 0) s.max_iter = o->max_iter;
    if (o->adaptive_rho >= 0) s.adaptive_rho = o->adaptive_rho;
    if (o->adaptive_rho_interval >= 0) s.adaptive_rho_interval = o->adaptive_rho_interval;
    if (o->check_termination >= 0) s.check_termination = o->check_termination;
    if (o->scaling >= 0) s.scaling = o->scaling;
    if (o->scaled_termination >= 0) s.scaled_termination = o->scaled_termination;
  }
  if (nthreads < 1) nthreads = 1;
  if (nthreads > B) nthreads = (int)(B > 0 ? B : 1);
  job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  c_int *perm = kkt_min_degree(n, m, P_colptr, P_rowidx, A_colptr, A_rowidx);   /* pattern is shared by the batch */
  double t0 = now_s();
  for (int t = 0; t < nthreads; ++t) {
    job_t *J = &jobs[t];
    J->n = n; J->m = m; J->nnzP = nnzP; J->nnzA = nnzA; J->perm = perm;
    J->P_colptr = P_colptr; J->P_rowidx = P_rowidx; J->A_colptr = A_colptr; J->A_rowidx = A_rowidx;
    J->P_val = P_val; J->q = q; J->A_val = A_val; J->l = l; J->u = u; J->warm_x = warm_x; J->warm_y = warm_y; J->s = s;
    J->x = x; J->y = y; J->obj = obj; J->pri_res = pri_res; J->dua_res = dua_res; J->setup_time = setup_time; J->solve_time = solve_time;
    J->wall_time = wall_time; J->status = status; J->iter = iter; J->rho_updates = rho_updates; J->exitflag = exitflag;
    J->dump = dump; J->dump_idx = dump_idx;
    J->begin = B * t / nthreads; J->end = B * (t + 1) / nthreads;
    if (nthreads == 1) worker(J); else pthread_create(&th[t], NULL, worker, J);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  double dt = now_s() - t0;
  free(jobs); free(th); free(perm);
  return dt;
}

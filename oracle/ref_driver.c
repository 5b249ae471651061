/* ORACLE / reference arm (test infrastructure, not product).
 *
 * Drives the reference's OWN solver binary — trajectory_planner/include/trajectory_planner/
 * third_party/lib/x86/libosqp.so (OSQP 0.6.2 + QDLDL + AMD; linked by the reference at
 * trajectory_planner/CMakeLists.txt:215) — through its C API exactly the way OsqpEigen::Solver does
 * for mpcPlanner::solveTraj (mpcPlanner.cpp:436-527):
 *     osqp_set_default_settings -> osqp_setup -> osqp_warm_start -> osqp_solve -> read -> osqp_cleanup
 * OSQP's source is not in the reference tree, so the binary is dlopen()ed from oracle/_ref/ where
 * oracle/Makefile places a copy (git-ignored).  Struct layouts below restate
 * third_party/osqp/types.h:21-289 under the build flags of third_party/osqp/osqp_configure.h:23-32
 * (PROFILING, DLONG, no DFLOAT, PRINTING, not EMBEDDED): c_int = long long, c_float = double.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) use this.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef long long c_int;   /* glob_opts.h:80 */
typedef double c_float;    /* glob_opts.h:87 */

typedef struct { c_int nzmax, m, n; c_int *p, *i; c_float *x; c_int nz; } r_csc;            /* types.h:21-29 */
typedef struct { c_float c; c_float *D, *E; c_float cinv; c_float *Dinv, *Einv; } r_scaling;/* types.h:42-49 */
typedef struct { c_float *x, *y; } r_solution;                                              /* types.h:55-58 */
typedef struct {                                                                            /* types.h:66-91 */
  c_int iter; char status[32]; c_int status_val; c_int status_polish;
  c_float obj_val, pri_res, dua_res;
  c_float setup_time, solve_time, update_time, polish_time, run_time;
  c_int rho_updates; c_float rho_estimate;
} r_info;
typedef struct { c_int n, m; r_csc *P, *A; c_float *q, *l, *u; } r_data;                    /* types.h:126-134 */
typedef struct {                                                                            /* types.h:139-176 */
  c_float rho, sigma; c_int scaling;
  c_int adaptive_rho, adaptive_rho_interval; c_float adaptive_rho_tolerance, adaptive_rho_fraction;
  c_int max_iter; c_float eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, alpha;
  int linsys_solver;            /* enum, 4 bytes + padding */
  c_float delta; c_int polish, polish_refine_iter, verbose;
  c_int scaled_termination, check_termination, warm_start;
  c_float time_limit;
} r_settings;
typedef struct {                                                                            /* types.h:182-289 */
  r_data *data; void *linsys_solver; void *pol;
  c_float *rho_vec, *rho_inv_vec; c_int *constr_type;
  c_float *x, *y, *z, *xz_tilde, *x_prev, *z_prev, *Ax, *Px, *Aty;
  c_float *delta_y, *Atdelta_y, *delta_x, *Pdelta_x, *Adelta_x;
  c_float *D_temp, *D_temp_A, *E_temp;
  r_settings *settings; r_scaling *scaling; r_solution *solution; r_info *info;
  void *timer; c_int first_run, clear_update_time, rho_update_from_solve; c_int summary_printed;
} r_workspace;

static void *g_lib;
static void (*p_set_default_settings)(r_settings *);
static c_int (*p_setup)(r_workspace **, const r_data *, const r_settings *);
static c_int (*p_warm_start)(r_workspace *, const c_float *, const c_float *);
static c_int (*p_solve)(r_workspace *);
static c_int (*p_cleanup)(r_workspace *);

/* Settings overrides passed from the caller; NaN / negative = keep OSQP default. */
typedef struct {
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, time_limit, adaptive_rho_tolerance;
  long long max_iter, adaptive_rho, adaptive_rho_interval, check_termination, scaling, warm_start, scaled_termination;
} ref_overrides;

int ref_open(const char *path) {
  if (g_lib) return 0;
  g_lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!g_lib) { fprintf(stderr, "ref_driver: dlopen(%s): %s\n", path, dlerror()); return 1; }
  p_set_default_settings = dlsym(g_lib, "osqp_set_default_settings");
  p_setup = dlsym(g_lib, "osqp_setup");
  p_warm_start = dlsym(g_lib, "osqp_warm_start");
  p_solve = dlsym(g_lib, "osqp_solve");
  p_cleanup = dlsym(g_lib, "osqp_cleanup");
  if (!p_set_default_settings || !p_setup || !p_warm_start || !p_solve || !p_cleanup) return 2;
  return 0;
}

void ref_default_settings(r_settings *s) { p_set_default_settings(s); }
int ref_sizeof_settings(void) { return (int)sizeof(r_settings); }

static void apply_overrides(r_settings *s, const ref_overrides *o) {
  if (!o) return;
  if (o->rho == o->rho) s->rho = o->rho;
  if (o->sigma == o->sigma) s->sigma = o->sigma;
  if (o->alpha == o->alpha) s->alpha = o->alpha;
  if (o->eps_abs == o->eps_abs) s->eps_abs = o->eps_abs;
  if (o->eps_rel == o->eps_rel) s->eps_rel = o->eps_rel;
  if (o->eps_prim_inf == o->eps_prim_inf) s->eps_prim_inf = o->eps_prim_inf;
  if (o->eps_dual_inf == o->eps_dual_inf) s->eps_dual_inf = o->eps_dual_inf;
  if (o->time_limit == o->time_limit) s->time_limit = o->time_limit;
  if (o->adaptive_rho_tolerance == o->adaptive_rho_tolerance) s->adaptive_rho_tolerance = o->adaptive_rho_tolerance;
  if (o->max_iter >= 0) s->max_iter = o->max_iter;
  if (o->adaptive_rho >= 0) s->adaptive_rho = o->adaptive_rho;
  if (o->adaptive_rho_interval >= 0) s->adaptive_rho_interval = o->adaptive_rho_interval;
  if (o->check_termination >= 0) s->check_termination = o->check_termination;
  if (o->scaling >= 0) s->scaling = o->scaling;
  if (o->warm_start >= 0) s->warm_start = o->warm_start;
  if (o->scaled_termination >= 0) s->scaled_termination = o->scaled_termination;
}

typedef struct {
  /* shared problem description */
  c_int n, m, nnzP, nnzA, B;
  const c_int *P_colptr, *P_rowidx, *A_colptr, *A_rowidx;
  const c_float *P_val, *q, *A_val, *l, *u, *warm_x, *warm_y;
  const ref_overrides *ov;
  /* outputs */
  c_float *x, *y, *obj, *pri_res, *dua_res, *setup_time, *solve_time, *wall_time;
  c_int *status, *iter, *rho_updates, *exitflag;
  /* optional internal dump of instance `dump_idx` (scaled iterates & scaling) */
  c_float *dump; c_int dump_idx;
  /* work split: a shared counter, every thread takes the next unsolved problem (a static split would leave the threads
   * that drew the 4000-iteration instances as the tail) */
  c_int *next;
} job_t;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void solve_one(const job_t *J, c_int b) {
  c_int n = J->n, m = J->m;
  double t0 = now_s();
  r_csc P = { J->nnzP, n, n, (c_int *)J->P_colptr, (c_int *)J->P_rowidx, (c_float *)(J->P_val + b * J->nnzP), -1 };
  r_csc A = { J->nnzA, m, n, (c_int *)J->A_colptr, (c_int *)J->A_rowidx, (c_float *)(J->A_val + b * J->nnzA), -1 };
  r_data d = { n, m, &P, &A, (c_float *)(J->q + b * n), (c_float *)(J->l + b * m), (c_float *)(J->u + b * m) };
  r_settings s; memset(&s, 0, sizeof s);
  p_set_default_settings(&s);
  s.verbose = 0;            /* mpcPlanner.cpp:440 */
  s.warm_start = 1;         /* mpcPlanner.cpp:441 */
  apply_overrides(&s, J->ov);
  r_workspace *w = NULL;
  c_int flag = p_setup(&w, &d, &s);
  if (flag || !w) {
    J->exitflag[b] = flag ? flag : 7; J->status[b] = -10; J->iter[b] = 0; J->rho_updates[b] = 0;
    J->obj[b] = 0; J->wall_time[b] = now_s() - t0; return;
  }
  if (J->warm_x) {  /* OsqpEigen::Solver::setWarmStart -> osqp_warm_start, Solver.tpp:216-244; dual = 0, mpcPlanner.cpp:487 */
    c_float *y0 = (c_float *)calloc((size_t)(m > 0 ? m : 1), sizeof(c_float));
    if (J->warm_y) memcpy(y0, J->warm_y + b * m, sizeof(c_float) * m);
    p_warm_start(w, J->warm_x + b * n, y0);
    free(y0);
  }
  c_int sf = p_solve(w);
  J->exitflag[b] = sf;
  J->status[b] = w->info->status_val; J->iter[b] = w->info->iter; J->rho_updates[b] = w->info->rho_updates;
  J->obj[b] = w->info->obj_val; J->pri_res[b] = w->info->pri_res; J->dua_res[b] = w->info->dua_res;
  J->setup_time[b] = w->info->setup_time; J->solve_time[b] = w->info->solve_time;
  memcpy(J->x + b * n, w->solution->x, sizeof(c_float) * n);
  if (J->y) memcpy(J->y + b * m, w->solution->y, sizeof(c_float) * m);
  if (J->dump && b == J->dump_idx) {
    /* layout: [c, rho, D(n), E(m), rho_vec(m), x(n), z(m), y(m)] — scaled internals after the solve */
    c_float *o = J->dump; *o++ = w->scaling->c; *o++ = w->settings->rho;
    memcpy(o, w->scaling->D, 8 * n); o += n; memcpy(o, w->scaling->E, 8 * m); o += m;
    memcpy(o, w->rho_vec, 8 * m); o += m;
    memcpy(o, w->x, 8 * n); o += n; memcpy(o, w->z, 8 * m); o += m; memcpy(o, w->y, 8 * m);
  }
  p_cleanup(w);
  J->wall_time[b] = now_s() - t0;
}

static void *worker(void *arg) {
  job_t *J = (job_t *)arg;
  for (;;) { c_int b = __atomic_fetch_add(J->next, 1, __ATOMIC_RELAXED); if (b >= J->B) break; solve_one(J, b); }
  return NULL;
}

/* Solve B problems sharing one CSC pattern, one problem per thread at a time (setup on the clock,
 * as the reference re-creates the solver per QP, mpcPlanner.cpp:436,527).  Returns total wall seconds. */
double ref_solve_batch(c_int n, c_int m, c_int nnzP, c_int nnzA, c_int B,
                       const c_int *P_colptr, const c_int *P_rowidx, const c_float *P_val, const c_float *q,
                       const c_int *A_colptr, const c_int *A_rowidx, const c_float *A_val,
                       const c_float *l, const c_float *u, const c_float *warm_x, const c_float *warm_y,
                       const ref_overrides *ov, int nthreads,
                       c_float *x, c_float *y, c_int *status, c_int *iter, c_int *rho_updates, c_int *exitflag,
                       c_float *obj, c_float *pri_res, c_float *dua_res, c_float *setup_time, c_float *solve_time,
                       c_float *wall_time, c_float *dump, c_int dump_idx) {
  if (!g_lib) return -1.0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > B) nthreads = (int)(B > 0 ? B : 1);
  job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  c_int next = 0;
  double t0 = now_s();
  for (int t = 0; t < nthreads; ++t) {
    job_t *J = &jobs[t];
    J->n = n; J->m = m; J->nnzP = nnzP; J->nnzA = nnzA; J->B = B;
    J->P_colptr = P_colptr; J->P_rowidx = P_rowidx; J->A_colptr = A_colptr; J->A_rowidx = A_rowidx;
    J->P_val = P_val; J->q = q; J->A_val = A_val; J->l = l; J->u = u; J->warm_x = warm_x; J->warm_y = warm_y; J->ov = ov;
    J->x = x; J->y = y; J->obj = obj; J->pri_res = pri_res; J->dua_res = dua_res;
    J->setup_time = setup_time; J->solve_time = solve_time; J->wall_time = wall_time;
    J->status = status; J->iter = iter; J->rho_updates = rho_updates; J->exitflag = exitflag;
    J->dump = dump; J->dump_idx = dump_idx;
    J->next = &next;
    if (nthreads == 1) worker(J); else pthread_create(&th[t], NULL, worker, J);
  }
  if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  double dt = now_s() - t0;
  free(jobs); free(th);
  return dt;
}

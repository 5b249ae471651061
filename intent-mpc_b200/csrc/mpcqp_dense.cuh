// mpcqp_dense.cuh — the generic (unstructured) solve path behind mpcqp_setup / mpcqp_solve: any convex QP in CSC form that
// does NOT have the mpcPlanner stage structure, i.e. the second in-tree consumer of the OsqpEigen::Solver boundary,
// polyTrajSolver (trajectory_planner/include/trajectory_planner/polyTrajSolver.cpp:14-37 constructs the three solvers,
// :162-239 setUpProblem / updateProblem, :848-900 solves them): minimum-snap coefficients of K path segments, n = 8K,
// m ~ 5K..7K equality rows (+ corridor boxes), n + m of a few hundred.
//
// One CTA per QP.  Same iterate sequence as OSQP 0.6.2 (third_party/osqp/auxil.h:21-154, constants.h:59-118): Ruiz
// equilibration, rho vector by constraint class, ADMM with relaxation, unscaled residual test, infeasibility certificates,
// rho adaptation with re-factorisation, OSQP_NAN fill.  Linear algebra B200-style instead of QDLDL's serial sparse
// triangular solves: the quasi-definite KKT matrix [P + sigma I, A'; A, -diag(1/rho)] (kkt.h:15-18) is held DENSE in
// L2-resident global memory, factored K = L D L' in place (right-looking, pivot column staged in shared memory, zero
// multipliers skipped), and the unit-triangular factor is inverted once per factorisation (one thread per column), so
// that every ADMM iteration's solve is two coalesced triangular mat-vecs (t = D^-1 L^-1 b, s = L^-T t) with one thread per
// row and two barriers, not 2(n+m) dependent steps.  Error growth is governed by cond(L), as for substitution.
//
// The file also compiles for the host (MPCQP_HOST_EMUL: one "thread", barriers vanish) so that the CPU test tier can check
// the logic against the reference binary; the shipped library has no host solve path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#ifdef MPCQP_HOST_EMUL
#define DQ_FN inline
#define DQ_SYNC() ((void)0)
#else
#define DQ_FN __device__ __forceinline__
#define DQ_SYNC() __syncthreads()
#endif
#define DQ_FOR(i, cnt) for (int i = tid; i < (cnt); i += nt)
// -DMPCQP_DENSE_PHASES: development build that prints clock64 totals per phase of one QP (tools/)
#if defined(MPCQP_DENSE_PHASES) && !defined(MPCQP_HOST_EMUL)
#define DQ_T(slot) do { DQ_SYNC(); const long long t_now = clock64(); ph[slot] += t_now - t_last; t_last = t_now; } while (0)
#else
#define DQ_T(slot) ((void)0)
#endif

namespace mpcqp_dense {

// third_party/osqp/constants.h:59-118
constexpr bool kRefine = true;
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3, kRhoTol = 1e-4, kMinScaling = 1e-4,
                 kMaxScaling = 1e4, kInfty = 1e30, kOsqpNan = 2143289344.0;   // OSQP_NAN is this NUMBER (constants.h:95-97)
enum { kSolved = 1, kSolvedInacc = 2, kPrimInfInacc = 3, kDualInfInacc = 4, kMaxIter = -2, kPrimInf = -3, kDualInf = -4,
       kNonCvx = -7, kUnsolved = -10 };

struct Settings {   // same layout as mpcqp::Settings (types.h:139-176 subset)
  double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, adaptive_rho_tolerance;
  int max_iter, scaling, adaptive_rho, adaptive_rho_interval, check_termination, warm_start;
};

struct Problem {    // pointers of ONE QP as the caller posed it (CSC, int64 indices like OSQP's c_int); nothing here is written
  int n, m;
  const int64_t *Pc, *Pi; const double* Px;   // P: upper-triangular CSC
  const int64_t *Ac, *Ai; const double* Ax;   // A: CSC, m x n
  const double *q0, *l0, *u0;
  const double *warm_x, *warm_y;   // user-space warm start or nullptr
  double* ws;       // ws_doubles(n, m), private to the CTA
  double *x, *y;    // out: [n], [m] (y may be nullptr)
  int32_t* info_i;  // out: status_val, iter, rho_updates
  double* info_d;   // out: obj_val, pri_res, dua_res
};

struct Batch {      // B QPs sharing one CSC pattern (polyTrajSolver's x / y / z problems; a set of candidate paths)
  int B, n, m;
  long long nnzP, nnzA;
  const int64_t *Pc, *Pi, *Ac, *Ai;
  const double *Px, *Ax, *q, *l, *u;          // [B][nnzP], [B][nnzA], [B][n], [B][m], [B][m]
  const double *warm_x, *warm_y;              // [B][n], [B][m] or nullptr
  double* ws; long long ws_stride;            // per resident CTA
  double *x, *y;                              // [B][n], [B][m] or nullptr
  int32_t* info_i; double* info_d;            // [B][3] each
};

// workspace: dense P (mirrored), A, A', the KKT factor (later L^-1) and L^-T, vectors
inline size_t ws_doubles(int n, int m) {
  const size_t N = (size_t)n + m;
  return 2 * N * N + (size_t)n * n + 2 * (size_t)n * m + 8 * N + 14 * (size_t)n + 18 * (size_t)m + 64;
}
constexpr int kMaxParts = 1;   // k-slices per row (groups of row threads); > 1 measured slower on B200 at 64 registers / thread
inline size_t smem_doubles(int n, int m) { return (3 + kMaxParts) * ((size_t)n + m) + 40; }

struct Solver {
  int tid, nt, n, m, N;
  Settings s;
  double *P, *A, *At, *q, *l, *u;
  double *L, *M1, *M2, *dinv, *tmpN;                    // factor, L^-1 (row i contiguous over threads), L^-T, 1/D
  double *D, *Dinv, *E, *Einv, *rho, *rho_inv, *ctype;
  double *x, *z, *y, *xp, *zp, *xt, *Axv, *Pxv, *Aty, *dy, *dx, *Atdy, *Pdx, *Adx, *tmpn, *tmpm;
  double *s_rhs, *s_t, *s_col, *s_red, *s_part;  // shared memory
  int RT, parts, rt, my_part;                    // row threads per group, groups, this thread's row lane and group
  double c, cinv;
  double pri_res, dua_res, obj;
  int status, rho_updates;
#if defined(MPCQP_DENSE_PHASES) && !defined(MPCQP_HOST_EMUL)
  long long ph[12], t_last;
#endif

  // ---- CTA-wide reductions; every thread gets the result --------------------------------------------------------------
  DQ_FN double blk_max(double v) {
#ifndef MPCQP_HOST_EMUL
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    DQ_SYNC();
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    DQ_SYNC();
    v = s_red[0];
    for (int w = 1; w < (nt >> 5); ++w) v = fmax(v, s_red[w]);
#endif
    return v;
  }
  DQ_FN double blk_sum(double v) {
#ifndef MPCQP_HOST_EMUL
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    DQ_SYNC();
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    DQ_SYNC();
    v = s_red[0];
    for (int w = 1; w < (nt >> 5); ++w) v += s_red[w];
#endif
    return v;
  }
  DQ_FN double norm_inf(const double* v, int cnt) { double a = 0; DQ_FOR(i, cnt) a = fmax(a, fabs(v[i])); return blk_max(a); }
  DQ_FN double scaled_norm_inf(const double* S, const double* v, int cnt) { double a = 0; DQ_FOR(i, cnt) a = fmax(a, fabs(S[i] * v[i])); return blk_max(a); }
  static DQ_FN double limit_scaling(double v) { v = v < kMinScaling ? 1.0 : v; return v > kMaxScaling ? kMaxScaling : v; }

  // ---- dense mat-vecs, one thread per output row, coalesced over the threads ----------------------------------------------
  DQ_FN void mv_A(const double* v, double* out) { DQ_FOR(i, m) { double a = 0; for (int j = 0; j < n; ++j) a += A[i + (size_t)m * j] * v[j]; out[i] = a; } }
  DQ_FN void mv_At(const double* v, double* out) { DQ_FOR(j, n) { double a = 0; for (int i = 0; i < m; ++i) a += At[j + (size_t)n * i] * v[i]; out[j] = a; } }
  DQ_FN void mv_P(const double* v, double* out) { DQ_FOR(j, n) { double a = 0; for (int i = 0; i < n; ++i) a += P[j + (size_t)n * i] * v[i]; out[j] = a; } }

  DQ_FN void carve(const Problem& pb, double* smem) {
    n = pb.n; m = pb.m; N = n + m;
    double* w = pb.ws; const size_t NN = (size_t)N * N;
    L = w; M1 = w; w += NN; M2 = w; w += NN;      // M1 (L^-1 for the first mat-vec) takes L's place once L^-1 is formed
    P = w; w += (size_t)n * n; A = w; w += (size_t)n * m; At = w; w += (size_t)n * m; q = w; w += n; l = w; w += m; u = w; w += m;
    dinv = w; w += N; xt = w; w += N; tmpN = w; w += N;
    D = w; w += n; Dinv = w; w += n; x = w; w += n; xp = w; w += n; Pxv = w; w += n; Aty = w; w += n; dx = w; w += n; Atdy = w; w += n;
    Pdx = w; w += n; tmpn = w; w += n;
    E = w; w += m; Einv = w; w += m; rho = w; w += m; rho_inv = w; w += m; ctype = w; w += m; z = w; w += m; zp = w; w += m; y = w; w += m;
    Axv = w; w += m; dy = w; w += m; Adx = w; w += m; tmpm = w; w += m;
    s_rhs = smem; s_t = smem + N; s_col = smem + 2 * (size_t)N; s_red = smem + 3 * (size_t)N; s_part = s_red + 40;
    const int NR = ((N + 31) / 32) * 32;
    RT = nt < NR ? nt : NR; parts = nt / RT; if (parts > kMaxParts) parts = kMaxParts;
    rt = tid % RT; my_part = tid / RT;
  }

  // ---- the caller's CSC -> dense scratch copies (P mirrored to full, A both ways), q, l, u ------------------------------------
  DQ_FN void densify(const Problem& pb) {
    for (size_t idx = tid; idx < (size_t)n * n; idx += nt) P[idx] = 0.0;
    for (size_t idx = tid; idx < (size_t)n * m; idx += nt) { A[idx] = 0.0; At[idx] = 0.0; }
    DQ_FOR(j, n) q[j] = pb.q0[j];
    DQ_FOR(i, m) { l[i] = pb.l0[i]; u[i] = pb.u0[i]; }
    DQ_SYNC();
    DQ_FOR(j, n) {     // thread j owns column j of P's upper triangle, its mirror image in row j, and column j of A
      for (long long t = pb.Pc[j]; t < pb.Pc[j + 1]; ++t) {
        const int i = (int)pb.Pi[t]; const double v = pb.Px[t];
        P[i + (size_t)n * j] += v;
        if (i != j) P[j + (size_t)n * i] += v;
      }
      for (long long t = pb.Ac[j]; t < pb.Ac[j + 1]; ++t) {
        const int i = (int)pb.Ai[t]; const double v = pb.Ax[t];
        A[i + (size_t)m * j] += v; At[j + (size_t)n * i] += v;
      }
    }
    DQ_SYNC();
  }

  // ---- scaling.h: scale_data (10 Ruiz passes + cost normalisation); same per-entry operation order as OSQP ------------------
  DQ_FN void scale_data() {
    DQ_FOR(j, n) D[j] = 1.0;
    DQ_FOR(i, m) E[i] = 1.0;
    c = 1.0;
    for (int pass = 0; pass < s.scaling; ++pass) {
      DQ_SYNC();
      DQ_FOR(j, n) {        // column inf-norms of [P A'; A 0], first n columns
        double a = 0;
        for (int i = 0; i < n; ++i) a = fmax(a, fabs(P[j + (size_t)n * i]));
        for (int i = 0; i < m; ++i) a = fmax(a, fabs(At[j + (size_t)n * i]));
        tmpn[j] = 1.0 / sqrt(limit_scaling(a));
      }
      DQ_FOR(i, m) {        // last m columns
        double a = 0;
        for (int j = 0; j < n; ++j) a = fmax(a, fabs(A[i + (size_t)m * j]));
        tmpm[i] = 1.0 / sqrt(limit_scaling(a));
      }
      DQ_SYNC();
      for (int idx = tid; idx < n * n; idx += nt) {
        const int i = idx % n, j = idx / n;
        P[idx] = (P[idx] * tmpn[i < j ? i : j]) * tmpn[i < j ? j : i];     // upper-triangle entry: row factor, then column factor
      }
      for (int idx = tid; idx < m * n; idx += nt) {
        const int i = idx % m, j = idx / m;
        A[idx] = (A[idx] * tmpm[i]) * tmpn[j];
      }
      for (int idx = tid; idx < m * n; idx += nt) {
        const int j = idx % n, i = idx / n;
        At[idx] = (At[idx] * tmpm[i]) * tmpn[j];
      }
      DQ_FOR(j, n) { q[j] *= tmpn[j]; D[j] *= tmpn[j]; }
      DQ_FOR(i, m) E[i] *= tmpm[i];
      DQ_SYNC();
      double part = 0, nq = 0;
      DQ_FOR(j, n) {
        double a = 0;
        for (int i = 0; i < n; ++i) a = fmax(a, fabs(P[j + (size_t)n * i]));
        part += a; nq = fmax(nq, fabs(q[j]));
      }
      double c_temp = blk_sum(part) / (double)n;
      nq = limit_scaling(blk_max(nq));
      if (nq > c_temp) c_temp = nq;
      c_temp = 1.0 / limit_scaling(c_temp);
      for (int idx = tid; idx < n * n; idx += nt) P[idx] *= c_temp;
      DQ_FOR(j, n) q[j] *= c_temp;
      c *= c_temp;
    }
    DQ_SYNC();
    cinv = 1.0 / c;
    DQ_FOR(j, n) Dinv[j] = 1.0 / D[j];
    DQ_FOR(i, m) { Einv[i] = 1.0 / E[i]; l[i] *= E[i]; u[i] *= E[i]; }
  }

  // ---- auxil.h: set_rho_vec (classes from the SCALED bounds) -------------------------------------------------------------
  DQ_FN void set_rho_vec() {
    s.rho = fmin(fmax(s.rho, kRhoMin), kRhoMax);
    DQ_FOR(i, m) {
      if (l[i] < -kInfty * kMinScaling && u[i] > kInfty * kMinScaling) { ctype[i] = -1; rho[i] = kRhoMin; }
      else if (u[i] - l[i] < kRhoTol) { ctype[i] = 1; rho[i] = kRhoEqOverIneq * s.rho; }
      else { ctype[i] = 0; rho[i] = s.rho; }
      rho_inv[i] = 1.0 / rho[i];
    }
  }

  // Row mat-vec helper: sum over k in [lo, hi) of M[i + N*k] * v[k] with eight independent partial sums (the loads are
  // what a thread waits for: L2 latency, not bandwidth, bounds these loops at one CTA per SM).
  DQ_FN double row_dot(const double* Mi, size_t ld, const double* v, int lo, int hi) {
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int k = lo;
    for (; k + 16 <= hi; k += 16) {      // sixteen matrix loads in flight (one L2 round trip per group is what a warp waits for)
      double mv[16];
#pragma unroll
      for (int t8 = 0; t8 < 16; ++t8) mv[t8] = Mi[ld * (k + t8)];
#pragma unroll
      for (int t8 = 0; t8 < 16; ++t8) a[t8 & 7] += mv[t8] * v[k + t8];
    }
    for (; k + 8 <= hi; k += 8) {
#pragma unroll
      for (int t8 = 0; t8 < 8; ++t8) a[t8] += Mi[ld * (k + t8)] * v[k + t8];
    }
    for (; k < hi; ++k) a[0] += Mi[ld * k] * v[k];
    return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
  // ---- KKT (kkt.h:15-18), dense lower triangle stored ROW-major (L[j + N*i] = K(i,j), i >= j: the threads of a warp sit on
  //      neighbouring columns), L D L' in place, then L^-1 ------------------------------------------------------------------
  DQ_FN void factor() {
    DQ_SYNC();
    for (int idx = tid; idx < N * N; idx += nt) {
      const int j = idx % N, i = idx / N;
      double v = 0.0;
      if (i >= j) {
        if (i < n) v = P[i + (size_t)n * j] + (i == j ? s.sigma : 0.0);
        else if (j < n) v = At[j + (size_t)n * (i - n)];
        else if (i == j) v = -rho_inv[i - n];
      }
      L[idx] = v;
    }
    DQ_T(8);
    // right-looking: step k subtracts c_i c_j / d_k from K(i,j), i >= j > k (c = the unscaled pivot column, staged in shared
    // memory).  Thread <-> column j; the threads of a warp walk the rows i in lockstep (coalesced), starting at the warp's
    // first diagonal; a column whose multiplier is zero (the KKT matrix is sparse until fill-in) is skipped.
    for (int k = 0; k < N; ++k) {
      DQ_SYNC();
      const int w = N - k - 1;
      DQ_FOR(ii, w) s_col[ii] = L[k + (size_t)N * (k + 1 + ii)];
      const double dk = L[k + (size_t)N * k];
      DQ_SYNC();
      for (int jj = tid; jj < w; jj += nt) {
        const double cj = s_col[jj];
        if (cj == 0.0) continue;
        const double f = cj / dk;
        double* colj = L + (k + 1 + jj) + (size_t)N * (k + 1);
#ifdef MPCQP_HOST_EMUL
        const int i0 = jj;
#else
        const int i0 = jj & ~31;
#endif
        for (int ii = i0; ii < w; ii += 16) {      // sixteen loads in flight per thread: the update is latency-bound otherwise
          double v[16];
#pragma unroll
          for (int t8 = 0; t8 < 16; ++t8) { const int r = ii + t8; v[t8] = (r < w) ? colj[(size_t)N * r] : 0.0; }
#pragma unroll
          for (int t8 = 0; t8 < 16; ++t8) { const int r = ii + t8; if (r >= jj && r < w) colj[(size_t)N * r] = v[t8] - s_col[r] * f; }
        }
      }
    }
    DQ_T(2);
    DQ_SYNC();
    DQ_FOR(k, N) dinv[k] = 1.0 / L[k + (size_t)N * k];
    DQ_SYNC();
    for (int idx = tid; idx < N * N; idx += nt) {
      const int j = idx % N, i = idx / N;
      if (i > j) L[idx] *= dinv[j];
    }
    DQ_SYNC();
    // column c of L^-1 by forward substitution, one thread per column; M2[c + N*i] = (L^-1)[i][c]; row i of L is contiguous
    DQ_FOR(cc, N) {
      for (int i = 0; i < N; ++i) {
        double a = (i == cc) ? 1.0 : 0.0;
        if (i > cc) a -= row_dot(M2 + cc, (size_t)N, L + (size_t)N * i, cc, i);     // eight partial sums: the chain is the critical path
        M2[cc + (size_t)N * i] = a;
      }
    }
    DQ_SYNC();
    for (int idx = tid; idx < N * N; idx += nt) {  // L is dead from here on: M1 overwrites it
      const int i = idx % N, k = idx / N;
      M1[idx] = M2[k + (size_t)N * i];           // M1[i + N*k] = (L^-1)[i][k]
    }
    DQ_SYNC();
    DQ_T(3);
  }

  // The CTA's threads form `parts` groups of RT row threads (carve): each group takes a contiguous slice of every row's k
  // range, the slices' sums meet in shared memory (s_part) and are added in a fixed order (deterministic).
  // b (shared) -> out (global): out = L^-T D^-1 L^-1 b; b is preserved
  DQ_FN void tri_solve(const double* b, double* out) {
    DQ_SYNC();
    if (my_part < parts) for (int i = rt; i < N; i += RT) {
      const int len = i + 1, lo = (int)((long long)len * my_part / parts), hi = (int)((long long)len * (my_part + 1) / parts);
      s_part[(size_t)my_part * N + i] = row_dot(M1 + i, (size_t)N, b, lo, hi);
    }
    DQ_SYNC();
    DQ_FOR(i, N) { double a = s_part[i]; for (int p = 1; p < parts; ++p) a += s_part[(size_t)p * N + i]; s_t[i] = a * dinv[i]; }
    DQ_SYNC();
    if (my_part < parts) for (int i = rt; i < N; i += RT) {
      const int len = N - i, lo = i + (int)((long long)len * my_part / parts), hi = i + (int)((long long)len * (my_part + 1) / parts);
      s_part[(size_t)my_part * N + i] = row_dot(M2 + i, (size_t)N, s_t, lo, hi);
    }
    DQ_SYNC();
    DQ_FOR(i, N) { double a = s_part[i]; for (int p = 1; p < parts; ++p) a += s_part[(size_t)p * N + i]; out[i] = a; }
    DQ_SYNC();
  }
  // s_rhs (shared) -> xt, with one step of iterative refinement against the KKT matrix in its original form (P, A, rho):
  // r = b - K s, s += K^-1 r.  The minimum-snap Hessians are ill-conditioned (entries 1 .. 1e5 per segment block); the
  // refined solve is at the accuracy of OSQP's own factorisation or better.  s_rhs is preserved.
  DQ_FN void kkt_solve() {
    DQ_T(6);
    tri_solve(s_rhs, xt);
    DQ_T(4);
    if (!kRefine) return;
    if (my_part < parts) for (int r = rt; r < N; r += RT) {     // K xt, row r, slice my_part of the n columns (+ of the m rows of A')
      const int lo = (int)((long long)n * my_part / parts), hi = (int)((long long)n * (my_part + 1) / parts);
      double a;
      if (r < n) {
        const int lo2 = (int)((long long)m * my_part / parts), hi2 = (int)((long long)m * (my_part + 1) / parts);
        a = row_dot(P + r, (size_t)n, xt, lo, hi) + row_dot(At + r, (size_t)n, xt + n, lo2, hi2);
      } else a = row_dot(A + (r - n), (size_t)m, xt, lo, hi);
      s_part[(size_t)my_part * N + r] = a;
    }
    DQ_SYNC();
    DQ_FOR(r, N) {
      double a = s_part[r]; for (int p = 1; p < parts; ++p) a += s_part[(size_t)p * N + r];
      a += (r < n) ? s.sigma * xt[r] : -rho_inv[r - n] * xt[r];
      s_col[r] = s_rhs[r] - a;
    }
    DQ_T(5);
    tri_solve(s_col, tmpN);
    DQ_FOR(i, N) xt[i] += tmpN[i];
    DQ_SYNC();
    DQ_T(4);
  }

  // ---- auxil.h:67-112, one ADMM iteration ------------------------------------------------------------------------------
  DQ_FN void iterate() {
    double* t;
    t = x; x = xp; xp = t;
    t = z; z = zp; zp = t;
    DQ_FOR(i, n) s_rhs[i] = s.sigma * xp[i] - q[i];
    DQ_FOR(i, m) s_rhs[n + i] = zp[i] - rho_inv[i] * y[i];
    kkt_solve();
    DQ_FOR(i, n) {
      x[i] = s.alpha * xt[i] + (1.0 - s.alpha) * xp[i];
      dx[i] = x[i] - xp[i];
    }
    DQ_FOR(i, m) {
      const double zt = s_rhs[n + i] + rho_inv[i] * xt[n + i];
      const double zr = s.alpha * zt + (1.0 - s.alpha) * zp[i];
      const double zn = fmin(fmax(zr + rho_inv[i] * y[i], l[i]), u[i]);      // proj.h: project
      z[i] = zn;
      dy[i] = rho[i] * (zr - zn);
      y[i] += dy[i];
    }
    DQ_SYNC();
  }

  // ---- update_info: unscaled residuals; the scaled ones stay in zp / xp for the rho estimate -------------------------------
  DQ_FN void update_info() {
    DQ_SYNC();
    if (m > 0) {
      mv_A(x, Axv);
      DQ_FOR(i, m) zp[i] = Axv[i] - z[i];
      pri_res = s.scaling ? scaled_norm_inf(Einv, zp, m) : norm_inf(zp, m);
    } else pri_res = 0.0;
    mv_P(x, Pxv);
    mv_At(y, Aty);
    DQ_FOR(j, n) xp[j] = (q[j] + Pxv[j]) + Aty[j];
    dua_res = s.scaling ? cinv * scaled_norm_inf(Dinv, xp, n) : norm_inf(xp, n);
  }

  DQ_FN bool is_primal_infeasible(double eps) {
    DQ_FOR(i, m) {
      if (u[i] > kInfty * kMinScaling) { if (l[i] < -kInfty * kMinScaling) dy[i] = 0.0; else dy[i] = fmin(dy[i], 0.0); }
      else if (l[i] < -kInfty * kMinScaling) dy[i] = fmax(dy[i], 0.0);
    }
    const double norm_dy = s.scaling ? scaled_norm_inf(E, dy, m) : norm_inf(dy, m);
    if (norm_dy > eps) {
      double part = 0;       // IEEE: +inf * 0 = NaN keeps the comparison false, as in OSQP
      DQ_FOR(i, m) part += u[i] * fmax(dy[i], 0.0) + l[i] * fmin(dy[i], 0.0);
      const double lhs = blk_sum(part);
      if (lhs < -eps * norm_dy) {
        DQ_SYNC();
        mv_At(dy, Atdy);
        const double nn = s.scaling ? scaled_norm_inf(Dinv, Atdy, n) : norm_inf(Atdy, n);
        return nn < eps * norm_dy;
      }
    }
    return false;
  }

  DQ_FN bool is_dual_infeasible(double eps) {
    const double norm_dx = s.scaling ? scaled_norm_inf(D, dx, n) : norm_inf(dx, n);
    const double cs = s.scaling ? c : 1.0;
    if (norm_dx > eps) {
      double part = 0;
      DQ_FOR(j, n) part += q[j] * dx[j];
      if (blk_sum(part) < -cs * eps * norm_dx) {
        DQ_SYNC();
        mv_P(dx, Pdx);
        const double np = s.scaling ? scaled_norm_inf(Dinv, Pdx, n) : norm_inf(Pdx, n);
        if (np < cs * eps * norm_dx) {
          mv_A(dx, Adx);
          double bad = 0;
          DQ_FOR(i, m) {
            const double a = s.scaling ? Adx[i] * Einv[i] : Adx[i];
            if ((u[i] < kInfty * kMinScaling && a > eps * norm_dx) || (l[i] > -kInfty * kMinScaling && a < -eps * norm_dx)) bad = 1.0;
          }
          return blk_max(bad) == 0.0;
        }
      }
    }
    return false;
  }

  DQ_FN bool check_termination(bool approximate) {
    double eps_abs = s.eps_abs, eps_rel = s.eps_rel, epi = s.eps_prim_inf, edi = s.eps_dual_inf;
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (pri_res > kInfty || dua_res > kInfty) { status = kNonCvx; obj = kOsqpNan; return true; }
    if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
    if (m == 0) prim_ok = true;
    else {
      const double a = s.scaling ? scaled_norm_inf(Einv, z, m) : norm_inf(z, m);
      const double b = s.scaling ? scaled_norm_inf(Einv, Axv, m) : norm_inf(Axv, m);
      if (pri_res < eps_abs + eps_rel * fmax(a, b)) prim_ok = true; else prim_inf = is_primal_infeasible(epi);
    }
    double t;
    if (s.scaling) {
      t = scaled_norm_inf(Dinv, q, n);
      t = fmax(t, scaled_norm_inf(Dinv, Aty, n));
      t = fmax(t, scaled_norm_inf(Dinv, Pxv, n));
      t *= cinv;
    } else { t = norm_inf(q, n); t = fmax(t, norm_inf(Aty, n)); t = fmax(t, norm_inf(Pxv, n)); }
    if (dua_res < eps_abs + eps_rel * t) dual_ok = true; else dual_inf = is_dual_infeasible(edi);
    if (prim_ok && dual_ok) { status = approximate ? kSolvedInacc : kSolved; return true; }
    if (prim_inf) { status = approximate ? kPrimInfInacc : kPrimInf; obj = kInfty; return true; }
    if (dual_inf) { status = approximate ? kDualInfInacc : kDualInf; obj = -kInfty; return true; }
    return false;
  }

  DQ_FN void adapt_rho() {
    double pr = norm_inf(zp, m), du = norm_inf(xp, n);
    const double pn = fmax(norm_inf(z, m), norm_inf(Axv, m));
    pr /= (pn + 1e-10);
    double dn = norm_inf(q, n); dn = fmax(dn, norm_inf(Aty, n)); dn = fmax(dn, norm_inf(Pxv, n));
    du /= (dn + 1e-10);
    const double rho_new = fmin(fmax(s.rho * sqrt(pr / (du + 1e-10)), kRhoMin), kRhoMax);
    if (rho_new > s.rho * s.adaptive_rho_tolerance || rho_new < s.rho / s.adaptive_rho_tolerance) {
      s.rho = rho_new;
      DQ_FOR(i, m) {
        if (ctype[i] == 0) { rho[i] = s.rho; rho_inv[i] = 1.0 / s.rho; }
        else if (ctype[i] == 1) { rho[i] = kRhoEqOverIneq * s.rho; rho_inv[i] = 1.0 / rho[i]; }
      }
      factor();
      rho_updates += 1;
    }
  }

  DQ_FN bool has_solution() const {
    return status != kPrimInf && status != kPrimInfInacc && status != kDualInf && status != kDualInfInacc && status != kNonCvx;
  }

  // ---- osqp_setup + osqp_warm_start + osqp_solve + store_solution for one QP ------------------------------------------------
  DQ_FN void run(const Problem& pb, const Settings& st, double* smem, int tid_, int nt_) {
    tid = tid_; nt = nt_; s = st;
    carve(pb, smem);
#if defined(MPCQP_DENSE_PHASES) && !defined(MPCQP_HOST_EMUL)
    for (int t = 0; t < 12; ++t) ph[t] = 0;
    t_last = clock64();
#endif
    densify(pb);
    DQ_T(0);
    if (s.scaling) scale_data();
    else { c = cinv = 1.0; DQ_FOR(j, n) D[j] = Dinv[j] = 1.0; DQ_FOR(i, m) E[i] = Einv[i] = 1.0; }
    DQ_SYNC();
    DQ_T(1);
    set_rho_vec();
    factor();
    // iterates: cold start or osqp_warm_start (x <- D^-1 x, y <- c E^-1 y, z <- A x)
    DQ_FOR(j, n) { x[j] = (pb.warm_x && s.warm_start) ? pb.warm_x[j] * Dinv[j] : 0.0; xp[j] = 0.0; }
    DQ_FOR(i, m) { y[i] = (pb.warm_y && s.warm_start) ? (pb.warm_y[i] * Einv[i]) * c : 0.0; zp[i] = 0.0; z[i] = 0.0; }
    DQ_SYNC();
    if (pb.warm_x && s.warm_start) mv_A(x, z);
    DQ_SYNC();
    status = kUnsolved; rho_updates = 0; obj = 0.0; pri_res = dua_res = 0.0;
    int iter, last = 0; bool can_check = false, done = false;
    for (iter = 1; iter <= s.max_iter; ++iter) {
      iterate();
      last = iter;
      can_check = s.check_termination && (iter % s.check_termination == 0);
      DQ_T(6);
      if (can_check) { update_info(); if (check_termination(false)) { done = true; break; } }
      DQ_T(7);
      if (s.adaptive_rho && s.adaptive_rho_interval && (iter % s.adaptive_rho_interval == 0)) {
        if (!can_check) update_info();
        adapt_rho();
      }
    }
    if (!done && !can_check) { update_info(); check_termination(false); }
    if (has_solution()) {
      DQ_SYNC();
      mv_P(x, Pxv);
      double part = 0;
      DQ_FOR(j, n) part += (0.5 * Pxv[j] + q[j]) * x[j];
      obj = blk_sum(part) * cinv;
    }
    if (status == kUnsolved) { if (!check_termination(true)) status = kMaxIter; }
    // store_solution (auxil.h:118)
    const bool ok = has_solution();
    DQ_FOR(j, n) pb.x[j] = ok ? (s.scaling ? D[j] * x[j] : x[j]) : kOsqpNan;
    if (pb.y) DQ_FOR(i, m) pb.y[i] = ok ? (s.scaling ? E[i] * y[i] * cinv : y[i]) : kOsqpNan;
#if defined(MPCQP_DENSE_PHASES) && !defined(MPCQP_HOST_EMUL)
    if (tid == 0 && blockIdx.x == 0) printf("phases N %d iters %d: densify %lld scale %lld kkt-fill %lld ldl %lld inverse %lld tri_solve %lld residual %lld iterate-rest %lld check %lld\n",
                                            N, last, ph[0], ph[1], ph[8], ph[2], ph[3], ph[4], ph[5], ph[6], ph[7]);
#endif
    if (tid == 0) {
      pb.info_i[0] = status; pb.info_i[1] = last; pb.info_i[2] = rho_updates;
      pb.info_d[0] = obj; pb.info_d[1] = pri_res; pb.info_d[2] = dua_res;
    }
  }
  DQ_FN void run_batch_item(const Batch& bt, int b, double* ws_cta, const Settings& st, double* smem, int tid_, int nt_) {
    Problem pb;
    pb.n = bt.n; pb.m = bt.m; pb.Pc = bt.Pc; pb.Pi = bt.Pi; pb.Ac = bt.Ac; pb.Ai = bt.Ai;
    pb.Px = bt.Px + (size_t)b * bt.nnzP; pb.Ax = bt.Ax + (size_t)b * bt.nnzA;
    pb.q0 = bt.q + (size_t)b * bt.n; pb.l0 = bt.l + (size_t)b * bt.m; pb.u0 = bt.u + (size_t)b * bt.m;
    pb.warm_x = bt.warm_x ? bt.warm_x + (size_t)b * bt.n : nullptr; pb.warm_y = bt.warm_y ? bt.warm_y + (size_t)b * bt.m : nullptr;
    pb.ws = ws_cta; pb.x = bt.x + (size_t)b * bt.n; pb.y = bt.y ? bt.y + (size_t)b * bt.m : nullptr;
    pb.info_i = bt.info_i + 3 * (size_t)b; pb.info_d = bt.info_d + 3 * (size_t)b;
    run(pb, st, smem, tid_, nt_);
  }
};

}  // namespace mpcqp_dense
